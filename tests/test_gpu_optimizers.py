"""Fused optimizers on the GPU vs torch.optim / in-file oracles (reference tests/L0/run_optimizers)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(dev, dtype=torch.float32):
    g = torch.Generator().manual_seed(0)
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(dev, dtype)) for s in [(4096, 64), (4096,), (278011,), (3, 5, 7)]]


def _drive(opt_a, opt_b, pa, pb, iters=7, set_none=False):
    g = torch.Generator().manual_seed(1)
    for _ in range(iters):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g)
            a.grad = gr.to(a.device, a.dtype)
            b.grad = gr.to(b.device, b.dtype)
        opt_a.step()
        opt_b.step()
        if set_none:
            opt_a.zero_grad()


@pytest.mark.parametrize("adam_w", [True, False])
def test_fused_adam_vs_torch(cuda_dev, adam_w):
    from apex_b200.optimizers import FusedAdam
    pa, pb = _params(cuda_dev), _params(cuda_dev)
    a = FusedAdam(pa, lr=5e-3, weight_decay=0.1, adam_w_mode=adam_w)
    b = (torch.optim.AdamW if adam_w else torch.optim.Adam)(pb, lr=5e-3, weight_decay=0.1)
    _drive(a, b, pa, pb, set_none=True)
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 1e-3


def test_fused_adam_table_is_cached(cuda_dev):
    from apex_b200.optimizers import FusedAdam
    pa = _params(cuda_dev)
    a = FusedAdam(pa, lr=1e-3)
    grads = [torch.randn_like(p) for p in pa]
    for _ in range(5):
        for p, g in zip(pa, grads):
            p.grad = g
        a.step()
    tb = a._cache._tables[(0, torch.float32)]
    assert tb.uploads == 1  # built once; the grad pointers never moved so nothing was re-uploaded


def test_fused_adam_follows_parameters_whose_storage_is_swapped(cuda_dev):
    """``p.data = ...`` after the first step (re-homing parameters into a flat buffer, ``.to()``): the cached device table must follow the
    new address instead of updating the old storage."""
    from apex_b200.optimizers import FusedAdam
    pa, pb = _params(cuda_dev), _params(cuda_dev)
    a, b = FusedAdam(pa, lr=5e-3, weight_decay=0.1), torch.optim.AdamW(pb, lr=5e-3, weight_decay=0.1)
    _drive(a, b, pa, pb, iters=2)
    flat = torch.cat([p.detach().reshape(-1) for p in pa])
    off = 0
    for p in pa:
        p.data = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    _drive(a, b, pa, pb, iters=3)
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 1e-3
    torch.testing.assert_close(flat, torch.cat([p.detach().reshape(-1) for p in pb]), rtol=1e-3, atol=1e-3)


def test_fused_adam_bf16_and_frozen_param(cuda_dev):
    from apex_b200.optimizers import FusedAdam
    pa = _params(cuda_dev, torch.bfloat16)
    a = FusedAdam(pa, lr=1e-2)
    for it in range(3):
        for i, p in enumerate(pa):
            p.grad = None if (i == 1 and it == 1) else torch.ones_like(p)
        before = pa[1].clone()
        a.step()
        if it == 1:
            assert torch.equal(before, pa[1])  # no grad -> untouched (table rebuilt)
    assert all(torch.isfinite(p.float()).all() for p in pa)


def test_fused_adam_capturable_with_grad_scaler(cuda_dev):
    from apex_b200.optimizers import FusedAdam
    pa, pb = _params(cuda_dev), _params(cuda_dev)
    a = FusedAdam(pa, lr=5e-3, capturable=True)
    b = torch.optim.AdamW(pb, lr=5e-3, weight_decay=0.0)
    scaler = torch.amp.GradScaler("cuda", init_scale=64.0)
    g = torch.Generator().manual_seed(1)
    for it in range(4):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=g).to(cuda_dev)
            x.grad = gr * 64.0
            y.grad = gr.clone()
        if it == 2:
            pa[0].grad[0, 0] = float("inf")  # this step must be skipped
            scaler._scale = torch.full((), 64.0, device=cuda_dev)
        scaler._lazy_init_scale_growth_tracker(cuda_dev) if scaler._scale is None else None
        scaler.step(a)
        scaler.update(64.0)
        if it != 2:
            b.step()
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 1e-3


def test_fused_sgd_adagrad_vs_torch(cuda_dev):
    from apex_b200.optimizers import FusedAdagrad, FusedSGD
    pa, pb = _params(cuda_dev), _params(cuda_dev)
    _drive(FusedSGD(pa, lr=0.05, momentum=0.9, weight_decay=0.01, nesterov=True),
           torch.optim.SGD(pb, lr=0.05, momentum=0.9, weight_decay=0.01, nesterov=True), pa, pb)
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 1e-3
    pa, pb = _params(cuda_dev), _params(cuda_dev)
    _drive(FusedAdagrad(pa, lr=0.05, weight_decay=0.01), torch.optim.Adagrad(pb, lr=0.05, weight_decay=0.01), pa, pb)
    for x, y in zip(pa, pb):
        assert (x - y).abs().max().item() <= 1e-3


def test_fused_lamb_and_novograd_vs_cpu_reference_path(cuda_dev):
    """The CPU path of the same optimizer class is plain PyTorch math: GPU kernels must agree with it."""
    from apex_b200.optimizers import FusedLAMB, FusedNovoGrad
    for Opt, kw in [(FusedLAMB, dict(lr=1e-2, weight_decay=0.01, max_grad_norm=1.0)), (FusedNovoGrad, dict(lr=1e-2, weight_decay=0.01))]:
        pa, pb = _params(cuda_dev), _params("cpu")
        _drive(Opt(pa, **kw), Opt(pb, **kw), pa, pb, iters=4)
        for x, y in zip(pa, pb):
            assert (x.cpu() - y).abs().max().item() <= 2e-3, Opt.__name__


def test_mixed_precision_lamb(cuda_dev):
    from apex_b200.optimizers import FusedMixedPrecisionLamb
    pa = _params(cuda_dev, torch.bfloat16)
    opt = FusedMixedPrecisionLamb(pa, lr=1e-2, reduced_precision_dtype=torch.bfloat16)
    ref0 = [p.clone() for p in pa]
    for _ in range(3):
        for p in pa:
            p.grad = torch.randn_like(p)
        opt.step()
    assert int(opt.param_groups[0]["step"].item()) == 3
    assert all(torch.isfinite(p.float()).all() for p in pa)
    assert any((p.float() - r.float()).abs().max() > 0 for p, r in zip(pa, ref0))
    sd = copy.deepcopy(opt.state_dict())
    opt.load_state_dict(sd)
    for p in pa:
        p.grad = torch.randn_like(p)
    opt.step()


@pytest.mark.parametrize("mode", ["PyTorchAdam", "ApexAdamW"])
def test_fused_adam_swa_single_launch(cuda_dev, mode):
    """openfold FusedAdamSWA: Adam + SWA average + bf16 compute copy (+ clip factor) in one multi-tensor launch vs a plain torch oracle."""
    from apex_b200.contrib.openfold.fused_adam_swa import AdamMathType, FusedAdamSWA
    torch.manual_seed(0)
    shapes = [(257, 33), (4099,), (64, 64)]
    p32 = [torch.nn.Parameter(torch.randn(s, device=cuda_dev)) for s in shapes]
    pbf = [torch.nn.Parameter(p.detach().bfloat16()) for p in p32]
    swa = [torch.nn.Parameter(p.detach().clone()) for p in p32]
    ref_p = [p.detach().clone() for p in p32]
    ref_m = [torch.zeros_like(p) for p in ref_p]
    ref_v = [torch.zeros_like(p) for p in ref_p]
    ref_swa = [p.clone() for p in ref_p]
    decay, lr, b1, b2, eps, wd = 0.9, 1e-2, 0.9, 0.999, 1e-8, 0.01
    opt = FusedAdamSWA(p32, pbf, swa, decay, lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd, adam_math_mode=AdamMathType[mode])
    for step in range(1, 5):
        clip = 0.5 if step == 3 else None
        for c in pbf:
            c.grad = torch.randn_like(c)
        grads = [c.grad.float() * (clip or 1.0) for c in pbf]
        opt.step(grad_clip_scale=clip)
        for i, g in enumerate(grads):
            p = ref_p[i]
            if mode == "PyTorchAdam":
                g = g + wd * p
            ref_m[i] = b1 * ref_m[i] + (1 - b1) * g
            ref_v[i] = b2 * ref_v[i] + (1 - b2) * g * g
            upd = (ref_m[i] / (1 - b1 ** step)) / ((ref_v[i] / (1 - b2 ** step)).sqrt() + eps)
            if mode == "ApexAdamW":
                upd = upd + wd * p
            ref_p[i] = p - lr * upd
            ref_swa[i] = ref_p[i].clone() if step == 1 else decay * ref_swa[i] + (1 - decay) * ref_p[i]
        for i in range(len(shapes)):
            torch.testing.assert_close(p32[i].detach(), ref_p[i], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(swa[i].detach(), ref_swa[i], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(pbf[i].detach().float(), ref_p[i].bfloat16().float(), rtol=0, atol=0)


def test_legacy_contrib_optimizers_on_gpu(cuda_dev):
    """Deprecated contrib optimizers on the multi-tensor kernels (reference apex/contrib/optimizers/{fused_adam,fp16_optimizer}.py): explicit
    ``grads`` / ``scale`` step, FP16_Optimizer master-weight round trip with a static loss scale, state_dict round trip."""
    from apex_b200.contrib.optimizers import FP16_Optimizer, FusedAdam
    from apex_b200.optimizers import FusedAdam as FA
    p = torch.nn.Parameter(torch.ones(1000, device=cuda_dev))
    o = FusedAdam([p], lr=0.1)
    o.step(grads=[torch.full((1000,), 4.0, device=cuda_dev)], scale=2.0)
    torch.testing.assert_close(p.detach(), torch.full((1000,), 0.9, device=cuda_dev))
    torch.manual_seed(0)
    m = torch.nn.Linear(64, 64).to(cuda_dev).half()
    ref = torch.nn.Linear(64, 64).to(cuda_dev)
    with torch.no_grad():
        ref.weight.copy_(m.weight.float())
        ref.bias.copy_(m.bias.float())
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-2)
    opt = FP16_Optimizer(FA(m.parameters(), lr=1e-2), static_loss_scale=8.0, verbose=False)
    for _ in range(3):
        x = torch.randn(16, 64, device=cuda_dev)
        opt.zero_grad()
        opt.backward(m(x.half()).float().pow(2).mean())
        opt.step()
        ropt.zero_grad()
        ref(x).pow(2).mean().backward()
        ropt.step()
    torch.testing.assert_close(m.weight.float(), ref.weight, atol=3e-3, rtol=3e-2)
    sd = opt.state_dict()
    opt.load_state_dict(sd)


def test_fused_adam_cuda_entry_points_on_gpu(cuda_dev):
    """reversible_adam + maybe_adam_undo + strided_check_finite + maybe_cast on the device (reference fused_adam_cuda.cpp:92-104)."""
    from apex_b200.contrib.optimizers import fused_adam_cuda as F
    torch.manual_seed(0)
    n = 4099
    for mode in (0, 1):
        p, m, v, g = (torch.randn(n, device=cuda_dev), torch.rand(n, device=cuda_dev) * 0.1, torch.rand(n, device=cuda_dev) * 0.1,
                      torch.randn(n, device=cuda_dev) * 4)
        p0, m0, v0 = p.clone(), m.clone(), v.clone()
        args = (1e-2, 0.9, 0.999, 1e-8, 2.0, 3, mode, 1, 0.01)
        copy = torch.empty(n, dtype=torch.bfloat16, device=cuda_dev)
        F.reversible_adam(p, copy, m, v, g, *args)
        assert not torch.equal(p, p0)
        torch.testing.assert_close(copy.float(), p, atol=2e-2, rtol=2e-2)
        F.maybe_adam_undo(torch.zeros(1, device=cuda_dev), p, m, v, g, *args)
        assert not torch.equal(p, p0)
        F.maybe_adam_undo(torch.ones(1, device=cuda_dev), p, m, v, g, *args)
        for a, b in ((p, p0), (m, m0), (v, v0)):
            torch.testing.assert_close(a, b, atol=1e-5, rtol=1e-4)
    flag, x = torch.zeros(1, device=cuda_dev), torch.ones(10, device=cuda_dev)
    x[4] = float("inf")
    F.strided_check_finite(flag, x, 2, 1)
    assert flag.item() == 1
    F.strided_check_finite(flag, x, 3, 1)
    assert flag.item() == 0
    src, dst = torch.randn(100, device=cuda_dev), torch.empty(100, dtype=torch.float16, device=cuda_dev)
    F.maybe_cast(torch.zeros(1, device=cuda_dev), src, dst)
    torch.testing.assert_close(dst.float(), src, atol=1e-3, rtol=1e-3)
