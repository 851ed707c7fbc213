"""Scaled softmax family, fused cross-entropy and RoPE kernels vs fp32 PyTorch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sk", [64, 128, 1000, 1024, 4096, 20000])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_scaled_masked_softmax(cuda_dev, sk, dtype):
    from apex_b200.transformer.functional import scaled_masked_softmax, scaled_softmax
    torch.manual_seed(0)
    b, h, sq = 2, 3, 17
    x = torch.randn(b, h, sq, sk, device=cuda_dev, dtype=dtype, requires_grad=True)
    mask = torch.rand(b, 1, sq, sk, device=cuda_dev) > 0.7
    mask[0, 0, 3] = True  # a fully masked row -> zeros
    xr = x.detach().float().requires_grad_(True)
    t = (xr * 0.7).masked_fill(mask, -10000.0)
    yr = torch.softmax(t, -1) * (~mask.all(-1, keepdim=True))
    y = scaled_masked_softmax(x, mask, 0.7)
    tol = 2e-3 if dtype == torch.float16 else 1e-2
    torch.testing.assert_close(y.float(), yr, atol=tol, rtol=tol)
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy.float())
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=tol, rtol=5 * tol)
    y2 = scaled_softmax(x.detach(), 1.3)
    torch.testing.assert_close(y2.float(), torch.softmax(x.detach().float() * 1.3, -1), atol=tol, rtol=tol)
    m1 = mask[:1]  # mask shared over the batch
    y3 = scaled_masked_softmax(x.detach(), m1, 0.7)
    t3 = (x.detach().float() * 0.7).masked_fill(m1, -10000.0)
    torch.testing.assert_close(y3.float(), torch.softmax(t3, -1) * (~m1.all(-1, keepdim=True)), atol=tol, rtol=tol)


@pytest.mark.parametrize("s", [8, 128, 1000, 2048])
def test_causal_softmax(cuda_dev, s):
    from apex_b200.transformer.functional import scaled_upper_triang_masked_softmax
    torch.manual_seed(0)
    x = torch.randn(5, s, s, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    xr = x.detach().float().requires_grad_(True)
    cm = torch.triu(torch.ones(s, s, dtype=torch.bool, device=cuda_dev), 1)
    yr = torch.softmax((xr * 0.5).masked_fill(cm, float("-inf")), -1)
    y = scaled_upper_triang_masked_softmax(x, 0.5)
    torch.testing.assert_close(y.float(), yr, atol=1e-2, rtol=1e-2)
    assert (y.masked_select(cm.expand_as(y)) == 0).all()
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy.float())
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=1e-2, rtol=5e-2)


@pytest.mark.parametrize("C", [32320, 1000, 50257])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("smoothing", [0.0, 0.1])
def test_xentropy(cuda_dev, C, dtype, smoothing):
    from apex_b200.contrib.xentropy import SoftmaxCrossEntropyLoss
    torch.manual_seed(0)
    N = 200
    x = (torch.randn(N, C, device=cuda_dev) * 2).to(dtype).requires_grad_(True)
    labels = torch.randint(0, C, (N,), device=cuda_dev)
    labels[::7] = 0  # padding rows
    xr = x.detach().float().requires_grad_(True)
    lp = torch.log_softmax(xr, -1)
    nll = -lp.gather(1, labels.view(-1, 1)).squeeze(1)
    ref = ((1 - smoothing) * nll - smoothing * lp.mean(-1)).masked_fill(labels == 0, 0.0)
    loss = SoftmaxCrossEntropyLoss.apply(x, labels, smoothing, 0, True)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    torch.testing.assert_close(loss, ref, atol=tol, rtol=tol)
    loss.sum().backward()
    ref.sum().backward()
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=tol, rtol=tol)


def _rope_ref(t, freqs):
    r = freqs.shape[-1]
    cos, sin = torch.cos(freqs), torch.sin(freqs)
    tr, tp = t[..., :r].float(), t[..., r:].float()
    x1, x2 = torch.chunk(tr, 2, -1)
    return torch.cat((tr * cos + torch.cat((-x2, x1), -1) * sin, tp), -1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("transpose", [False, True])
def test_rope_sbhd_and_cached(cuda_dev, dtype, transpose):
    from apex_b200.transformer.functional import fused_apply_rotary_pos_emb, fused_apply_rotary_pos_emb_cached
    torch.manual_seed(0)
    s, b, h, d, r = 37, 3, 5, 64, 48
    t = torch.randn(s, b, h, d, device=cuda_dev, dtype=dtype, requires_grad=True)
    freqs = torch.randn(s, 1, 1, r, device=cuda_dev)
    tr = t.detach().float().requires_grad_(True)
    ref = _rope_ref(tr, freqs)
    out = fused_apply_rotary_pos_emb(t, freqs, transpose)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    torch.testing.assert_close(out.float(), ref, atol=tol, rtol=tol)
    dy = torch.randn(s, b, h, d, device=cuda_dev, dtype=dtype)
    out.backward(dy)
    ref.backward(dy.float())
    torch.testing.assert_close(t.grad.float(), tr.grad, atol=tol, rtol=tol)
    out2 = fused_apply_rotary_pos_emb_cached(t.detach(), torch.cos(freqs).to(dtype), torch.sin(freqs).to(dtype), transpose)
    torch.testing.assert_close(out2.float(), ref.detach(), atol=max(tol, 2e-2 if dtype != torch.float32 else tol), rtol=2e-2)


def test_rope_thd_and_2d(cuda_dev):
    from apex_b200.transformer.functional import fused_apply_rotary_pos_emb_2d, fused_apply_rotary_pos_emb_thd
    torch.manual_seed(0)
    h, d = 4, 32
    lens = [5, 11, 1, 20]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), device=cuda_dev, dtype=torch.int32)
    T = sum(lens)
    t = torch.randn(T, h, d, device=cuda_dev, requires_grad=True)
    freqs = torch.randn(max(lens), 1, 1, d, device=cuda_dev)
    out = fused_apply_rotary_pos_emb_thd(t, cu, freqs)
    refs = []
    for i, L in enumerate(lens):
        seg = t.detach()[cu[i]:cu[i + 1]].unsqueeze(1)
        refs.append(_rope_ref(seg, freqs[:L]).squeeze(1))
    torch.testing.assert_close(out, torch.cat(refs), atol=1e-5, rtol=1e-5)
    out.sum().backward()
    assert torch.isfinite(t.grad).all()
    # 2-D
    b, ih, iw = 2, 6, 5
    x = torch.randn(b, ih * iw, h, d, device=cuda_dev, requires_grad=True)
    fh, fw = torch.randn(1, ih, 1, d // 2, device=cuda_dev), torch.randn(1, iw, 1, d // 2, device=cuda_dev)
    o = fused_apply_rotary_pos_emb_2d(x, ih, iw, torch.cos(fh), torch.sin(fh), torch.cos(fw), torch.sin(fw))
    x5 = x.detach().view(b, ih, iw, h, d)
    r1 = _rope_ref(x5[..., :d // 2], fh.view(1, ih, 1, 1, d // 2))
    r2 = _rope_ref(x5[..., d // 2:], fw.view(1, 1, iw, 1, d // 2))
    torch.testing.assert_close(o, torch.cat((r1, r2), -1).view(b, ih * iw, h, d), atol=1e-5, rtol=1e-5)
    # backward is the transpose of forward: <R x, y> == <x, R^T y>
    y = torch.randn_like(o)
    o.backward(y)
    x2 = torch.randn_like(x)
    o2 = fused_apply_rotary_pos_emb_2d(x2, ih, iw, torch.cos(fh), torch.sin(fh), torch.cos(fw), torch.sin(fw))
    torch.testing.assert_close((o2 * y).sum(), (x2 * x.grad).sum(), atol=1e-2, rtol=1e-3)
