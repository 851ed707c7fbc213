"""GPU numerics of the contrib kernels against PyTorch fp32 oracles: transducer joint / loss, focal loss, index_mul_2d,
clip_grad, multihead attention. Mirrors apex/contrib/test/{transducer,focal_loss,index_mul_2d,clip_grad,multihead_attn}."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lens(B, T, U, dev):
    g = torch.Generator().manual_seed(3)
    f_len = torch.randint(T // 2, T + 1, (B,), generator=g)
    y_len = torch.randint(U // 2, U, (B,), generator=g)
    f_len[0], y_len[-1] = T, U - 1
    return f_len.to(dev).int(), y_len.to(dev).int()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("pack", [False, True])
def test_transducer_joint(cuda_dev, dtype, relu, pack):
    from apex_b200.contrib.transducer import TransducerJoint
    from apex_b200.contrib.transducer.transducer import _TorchTransducerJoint
    B, T, U, H = 4, 13, 7, 72
    torch.manual_seed(0)
    f = torch.randn(B, T, H, device=cuda_dev, dtype=dtype, requires_grad=True)
    g = torch.randn(B, U, H, device=cuda_dev, dtype=dtype, requires_grad=True)
    f_len, y_len = _lens(B, T, U, cuda_dev)
    g_len = y_len + 1
    bo = torch.cumsum(f_len.long() * g_len.long(), 0)
    pb = int(bo[-1])
    out = TransducerJoint(pack_output=pack, relu=relu)(f, g, f_len, g_len, batch_offset=bo, packed_batch=pb)
    fr, gr = f.detach().float().requires_grad_(True), g.detach().float().requires_grad_(True)
    ref = _TorchTransducerJoint.forward(_TorchTransducerJoint(pack_output=pack, relu=relu), fr, gr, f_len, g_len, batch_offset=bo, packed_batch=pb)
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    torch.testing.assert_close(out.float(), ref, atol=tol, rtol=tol)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    out.backward(dy.to(dtype))
    torch.testing.assert_close(f.grad.float(), fr.grad, atol=tol * 20, rtol=tol * 20)
    torch.testing.assert_close(g.grad.float(), gr.grad, atol=tol * 20, rtol=tol * 20)


def test_transducer_joint_dropout(cuda_dev):
    from apex_b200.contrib.transducer import TransducerJoint
    B, T, U, H = 2, 9, 5, 64
    f = torch.randn(B, T, H, device=cuda_dev, requires_grad=True)
    g = torch.randn(B, U, H, device=cuda_dev, requires_grad=True)
    f_len = torch.tensor([9, 6], device=cuda_dev).int()
    g_len = torch.tensor([5, 3], device=cuda_dev).int()
    j = TransducerJoint(dropout=True, dropout_prob=0.25, probe_mask=True)
    out = j(f, g, f_len, g_len)
    mask = j.mask_probe[0]
    ref = (f.detach().unsqueeze(2) + g.detach().unsqueeze(1)) * mask / 0.75
    valid = (torch.arange(T, device=cuda_dev).view(1, T, 1) < f_len.view(B, 1, 1)) & (torch.arange(U, device=cuda_dev).view(1, 1, U) < g_len.view(B, 1, 1))
    torch.testing.assert_close(out, ref * valid.unsqueeze(-1), atol=1e-5, rtol=1e-5)
    keep = mask[valid].float().mean().item()
    assert 0.70 < keep < 0.80
    out.sum().backward()
    torch.testing.assert_close(f.grad, (mask.float() / 0.75 * valid.unsqueeze(-1)).sum(2), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("packed", [False, True])
def test_transducer_loss(cuda_dev, dtype, packed):
    from apex_b200.contrib.transducer import TransducerLoss
    from apex_b200.contrib.transducer.transducer import _TorchTransducerLoss
    B, T, U, V = 5, 17, 9, 37
    torch.manual_seed(1)
    f_len, y_len = _lens(B, T, U, cuda_dev)
    label = torch.randint(1, V, (B, U - 1), device=cuda_dev).int()
    x = torch.randn(B, T, U, V, device=cuda_dev, dtype=dtype)
    xr = x.detach().float().requires_grad_(True)
    ref = _TorchTransducerLoss.forward(_TorchTransducerLoss(), xr, label, f_len, y_len, 0)
    w = torch.rand(B, device=cuda_dev) + 0.5
    (ref * w).sum().backward()
    valid = (torch.arange(T, device=cuda_dev).view(1, T, 1) < f_len.view(B, 1, 1)) & (torch.arange(U, device=cuda_dev).view(1, 1, U) < (y_len + 1).view(B, 1, 1))
    if packed:
        bo = torch.cumsum(f_len.long() * (y_len.long() + 1), 0)
        xin = x[valid].clone().requires_grad_(True)
        loss = TransducerLoss(packed_input=True)(xin, label, f_len, y_len, 0, batch_offset=bo, max_f_len=T)
    else:
        xin = x.clone().requires_grad_(True)
        loss = TransducerLoss()(xin, label, f_len, y_len, 0)
    tol = 1e-4 if dtype == torch.float32 else 5e-2
    torch.testing.assert_close(loss.float(), ref.detach(), atol=tol, rtol=1e-4 if dtype == torch.float32 else 2e-3)
    (loss * w).sum().backward()
    gref = xr.grad * valid.unsqueeze(-1)
    got = xin.grad.float()
    gtol = 1e-4 if dtype == torch.float32 else 5e-3
    torch.testing.assert_close(got, gref[valid] if packed else gref, atol=gtol, rtol=gtol)


def test_focal_loss(cuda_dev):
    from apex_b200.contrib.focal_loss import focal_loss
    from apex_b200.contrib.focal_loss.focal_loss import _ref
    torch.manual_seed(0)
    N, C = 4096, 91
    x = torch.randn(N, C, device=cuda_dev, requires_grad=True)
    tgt = torch.randint(-2, 80, (N,), device=cuda_dev)
    npos = torch.tensor([123.0], device=cuda_dev)
    loss = focal_loss(x, tgt, npos, 80, 0.25, 2.0, 0.1)
    xr = x.detach().clone().requires_grad_(True)
    ref = _ref(xr, tgt, npos, 80, 0.25, 2.0, 0.1)
    torch.testing.assert_close(loss, ref.reshape(loss.shape), atol=1e-3, rtol=1e-4)
    loss.backward()
    ref.backward()
    torch.testing.assert_close(x.grad, xr.grad, atol=1e-6, rtol=1e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_index_mul_2d(cuda_dev, dtype):
    from apex_b200.contrib.index_mul_2d import index_mul_2d
    torch.manual_seed(0)
    in1 = torch.randn(500, 64, device=cuda_dev, dtype=dtype, requires_grad=True)
    in2 = torch.randn(6000, 64, device=cuda_dev, dtype=dtype, requires_grad=True)
    idx = torch.randint(0, 500, (6000,), device=cuda_dev)
    out = index_mul_2d(in1, in2, idx)
    a, b = in1.detach().float().requires_grad_(True), in2.detach().float().requires_grad_(True)
    ref = a[idx] * b
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    torch.testing.assert_close(out.float(), ref, atol=tol, rtol=tol)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    out.backward(dy.to(dtype))
    torch.testing.assert_close(in1.grad.float(), a.grad, atol=tol * 20, rtol=tol * 5)
    torch.testing.assert_close(in2.grad.float(), b.grad, atol=tol * 5, rtol=tol * 5)


def test_clip_grad(cuda_dev):
    from apex_b200.contrib.clip_grad import clip_grad_norm_
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(n, device=cuda_dev, dtype=dt)) for n, dt in [(1000, torch.float32), (37, torch.float16), (70000, torch.float32)]]
    rs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    for p, r in zip(ps, rs):
        p.grad = torch.randn_like(p)
        r.grad = p.grad.clone()
    n1 = clip_grad_norm_(ps, 0.5)
    n2 = torch.nn.utils.clip_grad_norm_(rs, 0.5)
    torch.testing.assert_close(n1.float().reshape(()), n2.float().reshape(()), atol=1e-3, rtol=1e-3)
    for p, r in zip(ps, rs):
        torch.testing.assert_close(p.grad.float(), r.grad.float(), atol=1e-3, rtol=2e-3)


def test_self_multihead_attn(cuda_dev):
    from apex_b200.contrib.multihead_attn import SelfMultiheadAttn
    torch.manual_seed(0)
    E, Hh, S, B = 128, 8, 24, 3
    m = SelfMultiheadAttn(E, Hh, dropout=0.0, bias=True).to(cuda_dev)
    ref = torch.nn.MultiheadAttention(E, Hh, dropout=0.0, bias=True).to(cuda_dev)
    from apex_b200.contrib.multihead_attn.multihead_attn import packed_to_blocked
    with torch.no_grad():   # packed projection rows are [head][q|k|v][head_dim] (the reference's layout); torch wants [q; k; v] blocks
        ref.in_proj_weight.copy_(packed_to_blocked(m.in_proj_weight, Hh))
        ref.in_proj_bias.copy_(packed_to_blocked(m.in_proj_bias, Hh))
        ref.out_proj.weight.copy_(m.out_proj_weight)
        ref.out_proj.bias.copy_(m.out_proj_bias)
    x = torch.randn(S, B, E, device=cuda_dev)
    out, _ = m(x, x, x, is_training=False)
    r, _ = ref(x, x, x, need_weights=False)
    torch.testing.assert_close(out, r, atol=2e-4, rtol=2e-4)


@pytest.mark.parametrize("world", [2, 4])
def test_peer_halo_exchange(cuda_dev, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.peer_halo_exchange_matches_allgather, world, "cuda", backend="nccl")
