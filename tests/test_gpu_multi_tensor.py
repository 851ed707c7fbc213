"""sm_100a multi-tensor kernels vs the plain-PyTorch fp32 reference (ops/reference.py). Sizes follow the reference test
vocabulary: 278011 (odd, unaligned tail), many small tensors, mixed dtypes, views at odd offsets."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [278011, 1, 7, 65536, 65537, 4096 * 33, 31]


def _mk(dev, dtype, sizes=SIZES, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [(torch.randn(n, generator=g) * scale).to(dev, dtype) for n in sizes]


def _clone(ls):
    return [[t.clone() for t in l] for l in ls]


def _close(a, b, tol):
    for x, y in zip(a, b):
        torch.testing.assert_close(x.float(), y.float(), atol=tol, rtol=tol)


@pytest.mark.parametrize("din,dout", [(torch.float32, torch.float32), (torch.float16, torch.float32), (torch.float32, torch.bfloat16)])
def test_scale(cuda_dev, din, dout):
    from apex_b200.ops import amp_C, reference as ref
    xs = _mk(cuda_dev, din)
    outs = [torch.empty(x.shape, device=cuda_dev, dtype=dout) for x in xs]
    noop = torch.zeros(1, dtype=torch.int, device=cuda_dev)
    amp_C.multi_tensor_scale(65536, noop, [xs, outs], 0.25)
    exp = [torch.empty_like(o) for o in outs]
    ref.multi_tensor_scale(None, [xs, exp], 0.25)
    _close(outs, exp, 1e-3)
    assert noop.item() == 0
    xs[3][17] = float("inf")
    amp_C.multi_tensor_scale(65536, noop, [xs, outs], 0.25)
    assert noop.item() == 1


def test_scale_unaligned_views(cuda_dev):
    from apex_b200.ops import amp_C
    base = torch.randn(100000, device=cuda_dev)
    xs = [base[1:1001], base[2003:9000]]
    outs = [torch.empty(1000, device=cuda_dev), torch.empty(6997, device=cuda_dev)]
    amp_C.multi_tensor_scale(65536, torch.zeros(1, dtype=torch.int, device=cuda_dev), [xs, outs], 2.0)
    _close(outs, [x * 2 for x in xs], 1e-6)


def test_axpby(cuda_dev):
    from apex_b200.ops import amp_C
    xs, ys = _mk(cuda_dev, torch.float16), _mk(cuda_dev, torch.float32, seed=1)
    outs = [torch.empty_like(y) for y in ys]
    noop = torch.zeros(1, dtype=torch.int, device=cuda_dev)
    amp_C.multi_tensor_axpby(65536, noop, [xs, ys, outs], 2.0, -0.5, -1)
    _close(outs, [2.0 * x.float() - 0.5 * y for x, y in zip(xs, ys)], 1e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_l2norm(cuda_dev, dtype):
    from apex_b200.ops import amp_C
    xs = _mk(cuda_dev, dtype)
    noop = torch.zeros(1, dtype=torch.int, device=cuda_dev)
    tot, per = amp_C.multi_tensor_l2norm(65536, noop, [xs], True)
    exp_per = torch.stack([x.float().norm() for x in xs])
    torch.testing.assert_close(per, exp_per, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(tot, exp_per.norm().reshape(1), rtol=1e-4, atol=1e-4)
    inv = torch.tensor([0.125], device=cuda_dev)
    tot2, _ = amp_C.multi_tensor_unscale_l2norm(65536, noop, [xs], inv, False)
    torch.testing.assert_close(tot2, tot * 0.125, rtol=1e-4, atol=1e-4)
    # l2norm + scale
    outs = [torch.empty_like(x) for x in xs]
    tot3, _ = amp_C.multi_tensor_l2norm_scale(65536, noop, [xs, outs], 0.5, False)
    torch.testing.assert_close(tot3, torch.stack([o.float().norm() for o in outs]).norm().reshape(1), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("gdt,pdt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float32)])
@pytest.mark.parametrize("mode", [0, 1])
def test_adam(cuda_dev, gdt, pdt, mode):
    from apex_b200.ops import amp_C, reference as ref
    g, p = _mk(cuda_dev, gdt, scale=0.1), _mk(cuda_dev, pdt, seed=1)
    m = [torch.zeros(x.shape, device=cuda_dev) for x in p]
    v = [torch.zeros(x.shape, device=cuda_dev) for x in p]
    a, b = [g, p, m, v], _clone([g, p, m, v])
    for step in range(1, 4):
        amp_C.multi_tensor_adam(65536, None, a, 1e-2, 0.9, 0.999, 1e-8, step, mode, 1, 0.05)
        ref.multi_tensor_adam(b, 1e-2, 0.9, 0.999, 1e-8, step, mode, 1, 0.05)
    tol = 2e-4 if pdt == torch.float32 else 2e-2  # division/sqrt ordering differs from the torch-op oracle by a few ulp
    _close(a[1], b[1], tol)
    _close(a[2], b[2], 1e-5)
    _close(a[3], b[3], 1e-5)


def test_adam_capturable_master_and_noop(cuda_dev):
    from apex_b200.ops import amp_C, reference as ref
    g, p = _mk(cuda_dev, torch.float16, scale=0.1), _mk(cuda_dev, torch.float16, seed=1)
    master = [x.float() for x in p]
    m = [torch.zeros(x.shape, device=cuda_dev) for x in p]
    v = [torch.zeros(x.shape, device=cuda_dev) for x in p]
    a, b = [g, p, m, v, master], _clone([g, p, m, v, master])
    lr = torch.tensor(1e-2, device=cuda_dev)
    step = torch.tensor([1], dtype=torch.int, device=cuda_dev)
    inv = torch.tensor([0.5], device=cuda_dev)
    noop = torch.zeros(1, dtype=torch.int, device=cuda_dev)
    amp_C.multi_tensor_adam_capturable_master(65536, noop, a, lr, 0.9, 0.999, 1e-8, step, 1, 1, 0.01, inv)
    ref.multi_tensor_adam_capturable(noop, b, lr, 0.9, 0.999, 1e-8, step, 1, 1, 0.01, inv)
    _close(a[4], b[4], 1e-5)
    _close(a[1], b[1], 2e-3)
    _close(a[0], b[0], 1e-3)
    before = [x.clone() for x in a[4]]
    noop.fill_(1)
    amp_C.multi_tensor_adam_capturable_master(65536, noop, a, lr, 0.9, 0.999, 1e-8, step, 1, 1, 0.01, inv)
    _close(a[4], before, 0.0)


def test_sgd_adagrad_novograd(cuda_dev):
    from apex_b200.ops import amp_C, reference as ref
    g, p = _mk(cuda_dev, torch.float32, scale=0.1), _mk(cuda_dev, torch.float32, seed=1)
    mom = [torch.zeros_like(x) for x in p]
    a, b = [g, p, mom], _clone([g, p, mom])
    for first in (True, False):
        amp_C.multi_tensor_sgd(65536, None, a, 0.01, 0.9, 0.0, 0.1, True, first, False, 1.0)
        ref.multi_tensor_sgd(None, b, 0.01, 0.9, 0.0, 0.1, True, first, False, 1.0)
    _close(a[1], b[1], 1e-5)
    _close(a[2], b[2], 1e-5)
    # fp16 grads, fp32 master + fp16 model copy
    g16 = _mk(cuda_dev, torch.float16, scale=0.1)
    model = [x.half() for x in p]
    a, b = [g16, [x.clone() for x in p], [torch.zeros_like(x) for x in p], model], None
    b = _clone(a)
    amp_C.multi_tensor_sgd(65536, None, a, 0.0, 0.9, 0.0, 0.1, False, True, False, 0.5)
    ref.multi_tensor_sgd(None, b, 0.0, 0.9, 0.0, 0.1, False, True, False, 0.5)
    _close(a[1], b[1], 1e-5)
    _close(a[3], b[3], 1e-3)
    # adagrad
    h = [torch.zeros_like(x) for x in p]
    a, b = [g, [x.clone() for x in p], h], None
    b = _clone(a)
    amp_C.multi_tensor_adagrad(65536, None, a, 0.01, 1e-10, 0, 0.01)
    ref.multi_tensor_adagrad(b, 0.01, 1e-10, 0, 0.01)
    _close(a[1], b[1], 1e-5)
    # novograd (both norm types)
    for norm_type in (2, 0):
        a = [g, [x.clone() for x in p], [torch.zeros_like(x) for x in p]]
        b = _clone(a)
        na = torch.stack([x.norm() if norm_type == 2 else x.abs().max() for x in g])
        nb = na.clone()
        for step in (1, 2):
            amp_C.multi_tensor_novograd(65536, None, a, na, 0.01, 0.95, 0.98, 1e-8, step, 1, 0.01, 1, 0, norm_type)
            ref.multi_tensor_novograd(b, nb, 0.01, 0.95, 0.98, 1e-8, step, 1, 0.01, 1, 0, norm_type)
        torch.testing.assert_close(na, nb, rtol=1e-4, atol=1e-5)
        _close(a[1], b[1], 1e-4)


@pytest.mark.parametrize("mode", [0, 1])
def test_lamb(cuda_dev, mode):
    from apex_b200.ops import amp_C, reference as ref
    g, p = _mk(cuda_dev, torch.float32, scale=0.5), _mk(cuda_dev, torch.float32, seed=1)
    m = [torch.zeros_like(x) for x in p]
    v = [torch.zeros_like(x) for x in p]
    a, b = [g, p, m, v], _clone([g, p, m, v])
    ggn = torch.stack([x.norm() for x in g]).norm().reshape(1)
    for step in (1, 2):
        amp_C.multi_tensor_lamb(65536, None, a, 1e-2, 0.9, 0.999, 1e-6, step, 1, 0.01, 1, mode, ggn, 1.0, False)
        ref.multi_tensor_lamb(b, 1e-2, 0.9, 0.999, 1e-6, step, 1, 0.01, 1, mode, ggn, 1.0, False)
    _close(a[1], b[1], 1e-4)
    _close(a[2], b[2], 1e-5)


def test_lamb_mp_with_model_copy(cuda_dev):
    from apex_b200.ops import amp_C, reference as ref
    g = _mk(cuda_dev, torch.bfloat16, scale=0.5)
    master = _mk(cuda_dev, torch.float32, seed=1)
    model = [x.bfloat16() for x in master]
    m = [torch.zeros_like(x) for x in master]
    v = [torch.zeros_like(x) for x in master]
    a = [g, master, m, v, model]
    b = _clone(a)
    lr = torch.tensor(1e-2, device=cuda_dev)
    step = torch.tensor([3], dtype=torch.int, device=cuda_dev)
    noop = torch.zeros(1, dtype=torch.int, device=cuda_dev)
    ggn = torch.stack([x.float().norm() for x in g]).norm().reshape(1)
    mx = torch.tensor([1.0], device=cuda_dev)
    inv = torch.tensor([1.0], device=cuda_dev)
    fi = torch.zeros(1, device=cuda_dev)
    amp_C.multi_tensor_lamb_mp(65536, noop, a, lr, 0.9, 0.999, 1e-6, step, 1, 0.01, 1, 1, ggn, mx, False, fi, inv)
    ref.multi_tensor_lamb_mp(noop, b, lr, 0.9, 0.999, 1e-6, step, 1, 0.01, 1, 1, ggn, mx, False, fi, inv)
    _close(a[1], b[1], 1e-4)
    _close(a[4], b[4], 1e-2)


def test_update_scale_hysteresis(cuda_dev):
    from apex_b200.ops import amp_C, reference as ref
    def run(fn, dev):
        scale = torch.tensor([1024.0], device=dev)
        gt = torch.zeros(1, dtype=torch.int, device=dev)
        ht = torch.tensor([2], dtype=torch.int, device=dev)
        out = []
        for inf in [0, 1, 0, 1, 1, 1, 0, 0, 0]:
            fn(scale, gt, ht, torch.tensor([float(inf)], device=dev), 2.0, 0.5, 2, 2)
            out.append((scale.item(), gt.item(), ht.item()))
        return out
    assert run(amp_C.update_scale_hysteresis, cuda_dev) == run(ref.update_scale_hysteresis, "cpu")


def test_cast_e5m2_roundtrip(cuda_dev):
    from apex_b200.ops import amp_C
    xs = _mk(cuda_dev, torch.float32)
    q = [torch.empty(x.shape, dtype=torch.float8_e5m2, device=cuda_dev) for x in xs]
    amp_C.multi_tensor_cast(65536, None, [xs, q])
    _close([t.float() for t in q], [x.to(torch.float8_e5m2).float() for x in xs], 0.0)


def test_ten_thousand_tensors_one_table(cuda_dev):
    from apex_b200.ops import amp_C
    g = torch.Generator().manual_seed(0)
    sizes = torch.randint(1, 3000, (10000,), generator=g).tolist()
    xs = [torch.randn(n, device=cuda_dev) for n in sizes]
    tb = amp_C.TensorTable([xs, [torch.empty_like(x) for x in xs]])
    amp_C.multi_tensor_scale(0, torch.zeros(1, dtype=torch.int, device=cuda_dev), tb, 3.0)
    outs = tb.slot(1)
    for i in (0, 17, 9999):
        torch.testing.assert_close(outs[i], xs[i] * 3.0)
    assert tb.uploads == 1
