"""Raw ``fast_multihead_attn`` entry points (contrib/multihead_attn/raw_ext.py) against plain autograd implementations that reuse the
returned dropout masks: forward intermediates (layouts of the reference, multihead_attn_frontend.cpp:573-607) and every backward."""
import pytest, torch, torch.nn.functional as F
from apex_b200.contrib.multihead_attn import raw_ext as X

T, S, B, H, E = 5, 7, 2, 4, 32
HD = E // H


def _attn(q, k, v, mask_fn, keep, p):
    s = torch.bmm(q, k.transpose(1, 2)) * HD ** -0.5
    pr = torch.softmax(mask_fn(s), -1)
    return torch.bmm(pr * keep.to(pr.dtype) / (1 - p), v).transpose(0, 1).reshape(q.shape[1], B, E), pr


def _mask_fn(kind, mask):
    if kind == "none":
        return lambda s: s
    if kind == "time":
        return lambda s: s.masked_fill(mask.bool()[None], float("-inf"))
    if kind == "pad":
        return lambda s: s.view(B, H, *s.shape[1:]).masked_fill(mask.bool()[:, None, None], float("-inf")).view(s.shape)
    return lambda s: (s.view(B, H, *s.shape[1:]) + mask[:, None, None]).view(s.shape)


def _masks(kind, Tq, Tk):
    if kind == "time":
        return torch.triu(torch.ones(Tq, Tk, dtype=torch.uint8), 1)
    if kind == "pad":
        m = torch.zeros(B, Tk, dtype=torch.uint8); m[1, -2:] = 1; return m
    if kind == "add":
        return torch.randn(B, Tk)
    return torch.tensor([])


@pytest.mark.parametrize("kind", ["none", "time", "pad"])
@pytest.mark.parametrize("bias", [False, True])
def test_self_attn_raw(kind, bias):
    torch.manual_seed(0)
    x = torch.randn(T, B, E, dtype=torch.double, requires_grad=True)
    wi, wo = torch.randn(3 * E, E, dtype=torch.double, requires_grad=True), torch.randn(E, E, dtype=torch.double, requires_grad=True)
    bi, bo = (torch.randn(3 * E, dtype=torch.double, requires_grad=True), torch.randn(E, dtype=torch.double, requires_grad=True)) if bias else (None, None)
    mask, p = _masks(kind, T, T), 0.25
    args = (kind != "none", kind == "time", True, H, x.detach(), wi.detach(), wo.detach())
    if bias:
        lin, probs, dropped, keep, ctx, out = X.self_attn_bias_forward(*args, bi.detach(), bo.detach(), mask, p)
    else:
        lin, probs, dropped, keep, ctx, out = X.self_attn_forward(*args, mask, p)
    assert lin.shape == (T, B, 3 * E) and probs.shape == (B * H, T, T) and keep.dtype == torch.uint8 and ctx.shape == (T, B * H, HD)
    l = F.linear(x, wi, bi).view(T, B * H, 3, HD)
    c, pr = _attn(l[:, :, 0].transpose(0, 1), l[:, :, 1].transpose(0, 1), l[:, :, 2].transpose(0, 1), _mask_fn(kind, mask), keep, p)
    ref = F.linear(c, wo, bo)
    torch.testing.assert_close(out, ref); torch.testing.assert_close(probs, pr)
    torch.testing.assert_close(dropped, pr * keep / (1 - p))
    g = torch.randn_like(ref)
    params = (x, wi, wo) + ((bi, bo) if bias else ())
    want = torch.autograd.grad(ref, params, g)
    fn = X.self_attn_bias_backward if bias else X.self_attn_backward
    got = fn(torch.tensor([H])[0], g, ctx, dropped, probs, lin, x.detach(), wi.detach(), wo.detach(), keep, torch.tensor([p])[0])
    assert len(got) == len(want)
    for a, r in zip(got, want):
        torch.testing.assert_close(a, r)


def test_self_attn_additive_mask_and_norm_add_raw():
    torch.manual_seed(1)
    x = torch.randn(T, B, E, dtype=torch.double, requires_grad=True)
    wi, wo = torch.randn(3 * E, E, dtype=torch.double, requires_grad=True), torch.randn(E, E, dtype=torch.double, requires_grad=True)
    bi, bo = torch.randn(3 * E, dtype=torch.double, requires_grad=True), torch.randn(E, dtype=torch.double, requires_grad=True)
    mask, p = _masks("add", T, T).double(), 0.2
    lin, bmm1, dropped, keep, ctx, out = X.self_attn_bias_additive_mask_forward(True, False, True, H, x.detach(), wi.detach(), wo.detach(), bi.detach(), bo.detach(), mask, p)
    l = F.linear(x, wi, bi).view(T, B * H, 3, HD)
    c, pr = _attn(l[:, :, 0].transpose(0, 1), l[:, :, 1].transpose(0, 1), l[:, :, 2].transpose(0, 1), _mask_fn("add", mask), keep, p)
    ref = F.linear(c, wo, bo)
    torch.testing.assert_close(out, ref)
    torch.testing.assert_close(bmm1, torch.bmm(l[:, :, 0].transpose(0, 1), l[:, :, 1].transpose(0, 1).transpose(1, 2)) * HD ** -0.5)
    g = torch.randn_like(ref)
    got = X.self_attn_bias_additive_mask_backward(H, g, ctx, dropped, bmm1, mask, lin, x.detach(), wi.detach(), wo.detach(), keep, p)
    for a, r in zip(got, torch.autograd.grad(ref, (x, wi, wo, bi, bo), g)):
        torch.testing.assert_close(a, r)
    # pre-LayerNorm + dropout + residual
    gam, bet = torch.randn(E, dtype=torch.double, requires_grad=True), torch.randn(E, dtype=torch.double, requires_grad=True)
    pm = _masks("pad", T, T)
    normed, mean, invvar, lin, probs, dropped, keep, ctx, add_keep, out = X.self_attn_norm_add_forward(True, False, True, H, x.detach(), gam.detach(), bet.detach(), wi.detach(), wo.detach(), pm, p)
    assert mean.shape == invvar.shape == (T * B,) and add_keep.shape == x.shape
    n = F.layer_norm(x, (E,), gam, bet, 1e-5)
    torch.testing.assert_close(normed, n)
    l = F.linear(n, wi).view(T, B * H, 3, HD)
    c, _ = _attn(l[:, :, 0].transpose(0, 1), l[:, :, 1].transpose(0, 1), l[:, :, 2].transpose(0, 1), _mask_fn("pad", pm), keep, p)
    ref = F.linear(c, wo) * add_keep / (1 - p) + x
    torch.testing.assert_close(out, ref)
    got = X.self_attn_norm_add_backward(H, g, ctx, dropped, probs, lin, normed, mean, invvar, x.detach(), gam.detach(), bet.detach(), wi.detach(), wo.detach(), keep, add_keep, p)
    for a, r in zip(got, torch.autograd.grad(ref, (x, gam, bet, wi, wo), g)):
        torch.testing.assert_close(a, r)


@pytest.mark.parametrize("kind", ["none", "time", "pad"])
def test_encdec_raw(kind):
    torch.manual_seed(2)
    xq = torch.randn(T, B, E, dtype=torch.double, requires_grad=True)
    xkv = torch.randn(S, B, E, dtype=torch.double, requires_grad=True)
    wq, wkv, wo = (torch.randn(n, E, dtype=torch.double, requires_grad=True) for n in (E, 2 * E, E))
    gam, bet = torch.randn(E, dtype=torch.double, requires_grad=True), torch.randn(E, dtype=torch.double, requires_grad=True)
    mask, p = _masks(kind, T, S), 0.3
    d = lambda *ts: [t.detach() for t in ts]
    lq, lkv, probs, dropped, keep, ctx, out = X.encdec_multihead_attn_forward(kind != "none", kind == "time", True, H, *d(xq, xkv, wq, wkv, wo), mask, p)
    assert lq.shape == (T, B, E) and lkv.shape == (S, B, 2 * E) and probs.shape == (B * H, T, S)

    def ref_fn(qin):
        q = F.linear(qin, wq).view(T, B * H, HD).transpose(0, 1)
        kv = F.linear(xkv, wkv).view(S, B * H, 2, HD)
        c, _ = _attn(q, kv[:, :, 0].transpose(0, 1), kv[:, :, 1].transpose(0, 1), _mask_fn(kind, mask), keep, p)
        return F.linear(c, wo)
    ref = ref_fn(xq)
    torch.testing.assert_close(out, ref)
    g = torch.randn_like(ref)
    got = X.encdec_multihead_attn_backward(H, g, ctx, dropped, probs, lq, lkv, *d(xq, xkv, wq, wkv, wo), keep, p)
    for a, r in zip(got, torch.autograd.grad(ref, (xq, xkv, wq, wkv, wo), g)):
        torch.testing.assert_close(a, r)
    torch.manual_seed(3)
    normed, mean, invvar, lq, lkv, probs, dropped, keep, ctx, add_keep, out = X.encdec_multihead_attn_norm_add_forward(
        kind != "none", kind == "time", True, H, *d(xq, xkv, gam, bet, wq, wkv, wo), mask, p)
    ref = ref_fn(F.layer_norm(xq, (E,), gam, bet, 1e-5)) * add_keep / (1 - p) + xq
    torch.testing.assert_close(out, ref)
    got = X.encdec_multihead_attn_norm_add_backward(H, g, ctx, dropped, probs, lq, lkv, normed, mean, invvar, *d(xq, xkv, gam, bet, wq, wkv, wo), keep, add_keep, p)
    for a, r in zip(got, torch.autograd.grad(ref, (xq, xkv, gam, bet, wq, wkv, wo), g)):
        torch.testing.assert_close(a, r)


def test_softmax_dropout_raw_and_extension_name():
    torch.manual_seed(4)
    s = torch.randn(B * H, T, S, dtype=torch.double, requires_grad=True)
    pm, am, p = _masks("pad", T, S), _masks("add", T, S).double(), 0.4
    for fwd, bwd, kind, m in ((X.mask_softmax_dropout_forward, X.mask_softmax_dropout_backward, "pad", pm),
                              (X.additive_mask_softmax_dropout_forward, X.additive_mask_softmax_dropout_backward, "add", am)):
        dropped, keep, probs = fwd(True, True, H, s.detach(), m, p)
        ref = torch.softmax(_mask_fn(kind, m)(s), -1) * keep / (1 - p)
        torch.testing.assert_close(dropped, ref)
        g = torch.randn_like(ref)
        got = bwd(True, H, g.clone(), probs, keep, m, p) if kind == "pad" else bwd(True, H, g.clone(), probs, keep, p)
        torch.testing.assert_close(got, torch.autograd.grad(ref, s, g)[0])
        dropped, keep, probs = fwd(False, False, H, s.detach(), torch.tensor([]), p)          # inference: no mask, no dropout
        torch.testing.assert_close(dropped, torch.softmax(s.detach(), -1)); assert bool(keep.all())
    import sys
    import apex_b200
    apex_b200.install_as_apex()
    import fast_multihead_attn
    assert len(X.ENTRY_POINTS) == 16 and all(hasattr(fast_multihead_attn, n) for n in X.ENTRY_POINTS)
