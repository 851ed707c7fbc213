"""tcgen05 attention kernels (csrc/fmha_{fwd,bwd}_sm100.cu) and the NVSwitch all-reduce against plain PyTorch fp32 references."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fmha():
    from apex_b200.contrib.fmha import kernels as X
    return X


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("seq", [128, 200, 517])
def test_fmha_fwd_fixed_length(cuda_dev, d, causal, seq):
    X = _fmha()
    torch.manual_seed(0)
    b, h = 3, 4
    qkv = torch.randn(b * seq, 3, h, d, device=cuda_dev, dtype=torch.bfloat16)
    out, lse = X.fmha_fwd(qkv[:, 0], qkv[:, 1], qkv[:, 2], batch=b, causal=causal, return_lse=True)
    q, k, v = (qkv[:, i].view(b, seq, h, d).transpose(1, 2).float() for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal)
    torch.testing.assert_close(out.view(b, seq, h, d).transpose(1, 2).float(), ref, atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("pattern", ["ramp", "spikes", "decay"])
def test_fmha_fwd_online_softmax_rescaling(cuda_dev, d, causal, pattern):
    """The forward is a single pass with LAZY rescaling of the TMEM accumulators (the exponent offset moves only when a tile's maximum
    exceeds it by 2^8). Random N(0, 1) inputs almost never trigger that after the first tile, so these inputs force it: key norms that
    grow along the sequence (a rescale at most tiles), isolated huge keys (rescale in one half of a tile only), and norms that shrink
    (never rescale: stale, too-large offsets must still give exact results, and the log-sum-exp must match)."""
    X = _fmha()
    torch.manual_seed(1)
    b, h, seq = 2, 3, 1000
    q = torch.randn(b, seq, h, d, device=cuda_dev)
    k = torch.randn(b, seq, h, d, device=cuda_dev)
    v = torch.randn(b, seq, h, d, device=cuda_dev)
    pos = torch.arange(seq, device=cuda_dev, dtype=torch.float32).view(1, seq, 1, 1) / seq
    if pattern == "ramp":
        k = k * (0.2 + 9.0 * pos)
    elif pattern == "decay":
        k = k * (9.0 - 8.8 * pos)
    else:
        k[:, 70::129] *= 25.0        # one key per tile, alternating between the two 64-key halves
    q, k, v = (t.to(torch.bfloat16) for t in (q, k, v))
    qf, kf, vf = (t.reshape(b * seq, h, d) for t in (q, k, v))
    out, lse = X.fmha_fwd(qf, kf, vf, batch=b, causal=causal, return_lse=True)
    qr, kr, vr = (t.transpose(1, 2).float() for t in (q, k, v))
    sc = qr @ kr.transpose(2, 3) / d ** 0.5
    if causal:
        sc = sc.masked_fill(torch.ones(seq, seq, device=cuda_dev, dtype=torch.bool).triu(1), float("-inf"))
    ref = torch.softmax(sc, -1) @ vr
    torch.testing.assert_close(out.view(b, seq, h, d).transpose(1, 2).float(), ref, atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(lse.view(b, seq, h).transpose(1, 2), torch.logsumexp(sc, -1), atol=2e-2, rtol=2e-3)


def test_fmha_fwd_varlen(cuda_dev):
    X = _fmha()
    torch.manual_seed(0)
    h, d = 4, 64
    lens = [5, 130, 1, 300]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=cuda_dev)
    qkv = torch.randn(sum(lens), 3, h, d, device=cuda_dev, dtype=torch.float16)
    out = X.fmha_fwd(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu_seqlens_q=cu, max_seqlen_q=max(lens))
    for i, n in enumerate(lens):
        s = int(cu[i])
        q, k, v = (qkv[s:s + n, j].transpose(0, 1).float() for j in range(3))
        ref = torch.softmax(q @ k.transpose(1, 2) / d ** 0.5, -1) @ v
        torch.testing.assert_close(out[s:s + n].transpose(0, 1).float(), ref, atol=2e-2, rtol=2e-2)



@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("seq", [128, 200, 333])
def test_fmha_bwd_fixed_length(cuda_dev, d, causal, seq):
    X = _fmha()
    torch.manual_seed(0)
    b, h = 2, 3
    qkv = torch.randn(b * seq, 3, h, d, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    out = X.FmhaFunc.apply(qkv[:, 0], qkv[:, 1], qkv[:, 2], None, None, None, None, b, causal, None)
    dout = torch.randn_like(out)
    (grad,) = torch.autograd.grad(out, qkv, dout)
    ref_in = qkv.detach().float().requires_grad_()
    q, k, v = (ref_in[:, i].view(b, seq, h, d).transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(b * seq, h, d)
    (ref_grad,) = torch.autograd.grad(ref, ref_in, dout.float())
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(grad.float(), ref_grad, atol=5e-2, rtol=5e-2)


def test_fmha_bwd_varlen(cuda_dev):
    X = _fmha()
    torch.manual_seed(0)
    h, d = 2, 64
    lens = [5, 130, 1, 300]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=cuda_dev)
    qkv = torch.randn(sum(lens), 3, h, d, device=cuda_dev, dtype=torch.float16, requires_grad=True)
    out = X.FmhaFunc.apply(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu, cu, max(lens), max(lens), None, False, None)
    dout = torch.randn_like(out)
    (grad,) = torch.autograd.grad(out, qkv, dout)
    ref_in = qkv.detach().float().requires_grad_()
    outs = []
    for i, n in enumerate(lens):
        s = int(cu[i])
        q, k, v = (ref_in[s:s + n, j].transpose(0, 1) for j in range(3))
        outs.append((torch.softmax(q @ k.transpose(1, 2) / d ** 0.5, -1) @ v).transpose(0, 1))
    (ref_grad,) = torch.autograd.grad(torch.cat(outs), ref_in, dout.float())
    torch.testing.assert_close(grad.float(), ref_grad, atol=5e-2, rtol=5e-2)


@pytest.mark.parametrize("world", [2, 8])
def test_nvls_allreduce(cuda_dev, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from apex_b200.parallel import nvls_allreduce as N
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.nvls_allreduce_matches_nccl, world, "cuda", backend="nccl")


def _ref_attention(q, k, v, scale, causal=False, key_bias=None, keep=None, p=0.0):
    """fp32 reference on [b, h, s, d] tensors; key_bias [b, sk]; keep: bool [b, h, sq, sk] dropout mask."""
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    if key_bias is not None:
        s = s + key_bias[:, None, None, :]
    if causal:
        sq, sk = s.shape[-2:]
        s = s.masked_fill(torch.ones(sq, sk, dtype=torch.bool, device=s.device).triu(1 + sk - sq), float("-inf"))
    pr = torch.softmax(s, -1)
    if keep is not None:
        pr = pr * keep / (1.0 - p)
    return torch.matmul(pr, v)


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("additive", [False, True])
def test_fmha_key_bias_fwd_bwd(cuda_dev, d, additive):
    X = _fmha()
    torch.manual_seed(1)
    b, h, sq, sk = 3, 2, 150, 217
    q = torch.randn(b * sq, h, d, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(b * sk, h, d, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(b * sk, h, d, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    pad = torch.zeros(b, sk, dtype=torch.bool, device=cuda_dev)
    pad[0, 200:] = True
    pad[2, 17:] = True
    bias = torch.randn(b, sk, device=cuda_dev) if additive else torch.zeros(b, sk, device=cuda_dev).masked_fill_(pad, float("-inf"))
    out = X.FmhaFunc.apply(q, k, v, None, None, None, None, b, False, None, bias, 0.0)
    dout = torch.randn_like(out)
    gq, gk, gv = torch.autograd.grad(out, (q, k, v), dout)
    qf, kf, vf = (t.detach().float().requires_grad_() for t in (q, k, v))
    r = lambda t, s: t.view(b, s, h, d).transpose(1, 2)  # noqa: E731
    ref = _ref_attention(r(qf, sq), r(kf, sk), r(vf, sk), d ** -0.5, key_bias=bias).transpose(1, 2).reshape(b * sq, h, d)
    rq, rk, rv = torch.autograd.grad(ref, (qf, kf, vf), dout.float())
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    for g, rg in ((gq, rq), (gk, rk), (gv, rv)):
        torch.testing.assert_close(g.float(), rg, atol=5e-2, rtol=5e-2)


@pytest.mark.parametrize("d,causal", [(64, False), (128, True)])
def test_fmha_dropout_matches_the_philox_reference_mask(cuda_dev, d, causal):
    X = _fmha()
    torch.manual_seed(2)
    b, h, s, p = 2, 3, 200, 0.2
    qkv = torch.randn(b * s, 3, h, d, device=cuda_dev, dtype=torch.float16, requires_grad=True)
    gen = torch.cuda.default_generators[cuda_dev.index or 0]
    philox = (gen.initial_seed() & (2**63 - 1), gen.get_offset())
    out = X.FmhaFunc.apply(qkv[:, 0], qkv[:, 1], qkv[:, 2], None, None, None, None, b, causal, None, None, p)
    assert gen.get_offset() > philox[1]
    dout = torch.randn_like(out)
    (grad,) = torch.autograd.grad(out, qkv, dout)
    keep = X.dropout_keep_mask(b, h, s, s, p, philox, device=cuda_dev)
    assert abs(keep.float().mean().item() - (1 - p)) < 0.02
    ref_in = qkv.detach().float().requires_grad_()
    q, k, v = (ref_in[:, i].view(b, s, h, d).transpose(1, 2) for i in range(3))
    ref = _ref_attention(q, k, v, d ** -0.5, causal=causal, keep=keep, p=p).transpose(1, 2).reshape(b * s, h, d)
    (ref_grad,) = torch.autograd.grad(ref, ref_in, dout.float())
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(grad.float(), ref_grad, atol=5e-2, rtol=5e-2)


def test_fmha_module_default_route_is_the_kernel(cuda_dev):
    """contrib.fmha.FMHA on fp16, d = 64: must launch ab_fmha_fwd / ab_fmha_bwd (no SDPA), varlen + dropout in training."""
    from types import SimpleNamespace
    from apex_b200 import _lib
    from apex_b200.contrib.fmha import FMHA

    torch.manual_seed(0)
    m = FMHA(SimpleNamespace(attention_probs_dropout_prob=0.1, num_attention_heads=4, hidden_size=256))
    lens = [37, 128, 300]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=cuda_dev)
    qkv = torch.randn(sum(lens), 3 * 256, device=cuda_dev, dtype=torch.float16, requires_grad=True)
    seen = []
    real = _lib.fn

    def spy(name):
        seen.append(name)
        return real(name)

    _lib.fn = spy
    try:
        out = m(qkv, cu, max(lens), is_training=True)
        out.float().sum().backward()
    finally:
        _lib.fn = real
    assert "ab_fmha_fwd" in seen and "ab_fmha_bwd" in seen and torch.isfinite(qkv.grad.float()).all()
    ev = m(qkv.detach(), cu, max(lens), is_training=False)
    ref = []
    x = qkv.detach().view(-1, 3, 4, 64).float()
    for i, n in enumerate(lens):
        s0 = int(cu[i])
        q, k, v = (x[s0:s0 + n, j].transpose(0, 1) for j in range(3))
        ref.append((torch.softmax(q @ k.transpose(1, 2) / 8.0, -1) @ v).transpose(0, 1).reshape(n, 256))
    torch.testing.assert_close(ev.float(), torch.cat(ref), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("mask", ["none", "padding", "additive", "causal"])
def test_self_multihead_attn_kernel_route_matches_generic(cuda_dev, mask, monkeypatch):
    from apex_b200.contrib.fmha import kernels as K
    from apex_b200.contrib.multihead_attn import SelfMultiheadAttn

    torch.manual_seed(0)
    t, b, e, heads = 130, 3, 256, 4
    mha = SelfMultiheadAttn(e, heads, dropout=0.0, bias=True, impl="fast", mask_additive=(mask == "additive")).to(cuda_dev, torch.bfloat16)
    x = torch.randn(t, b, e, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    kw = {}
    if mask == "padding":
        kpm = torch.zeros(b, t, dtype=torch.bool, device=cuda_dev)
        kpm[1, 100:] = True
        kw["key_padding_mask"] = kpm
    elif mask == "additive":
        kw["key_padding_mask"] = torch.randn(b, t, device=cuda_dev, dtype=torch.bfloat16)
    elif mask == "causal":
        kw["attn_mask"] = torch.ones(t, t, dtype=torch.bool, device=cuda_dev).triu(1)
    out, _ = mha(x, is_training=True, **kw)
    (g,) = torch.autograd.grad(out, x, torch.ones_like(out))
    monkeypatch.setattr(K, "supported", lambda *a: False)
    ref, _ = mha(x, is_training=True, **kw)
    (rg,) = torch.autograd.grad(ref, x, torch.ones_like(ref))
    torch.testing.assert_close(out.float(), ref.float(), atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(g.float(), rg.float(), atol=6e-2, rtol=6e-2)
