"""The CPU reference implementations are the oracles the GPU kernel tests compare against. Here the oracles themselves are checked against
independent formulations (so that a kernel and its reference cannot agree on the same mistake): RoPE against explicit pair rotations, the
transducer loss against an enumeration of every alignment."""
import itertools
import math
import random

import pytest
import torch


def _rotate_pairs(t, freqs):
    """Pairs (i, i + r/2) of the first r features rotate by the angle freqs[..., i] (the second half by freqs[..., i + r/2])."""
    r = freqs.shape[-1]
    x1, x2, rest = t[..., : r // 2].double(), t[..., r // 2: r].double(), t[..., r:].double()
    a1, a2 = freqs[..., : r // 2].double(), freqs[..., r // 2:].double()
    return torch.cat([x1 * a1.cos() - x2 * a1.sin(), x2 * a2.cos() + x1 * a2.sin(), rest], -1).to(t.dtype)


@pytest.mark.parametrize("seed", range(12))
def test_rope_references_against_explicit_pair_rotation(seed):
    from apex_b200.transformer.functional import fused_rope as R
    rng = random.Random(seed)
    torch.manual_seed(seed)
    s, b, h, d = rng.randint(1, 9), rng.randint(1, 3), rng.randint(1, 4), rng.choice([8, 16, 32])
    r = rng.choice([d, d // 2])
    t = torch.randn(s, b, h, d, requires_grad=True)
    freqs = torch.randn(s, 1, 1, r)
    out = R.fused_apply_rotary_pos_emb(t, freqs)
    want = _rotate_pairs(t.detach(), freqs)
    torch.testing.assert_close(out, want, atol=1e-5, rtol=1e-5)
    g = torch.randn_like(out)
    t2 = t.detach().clone().requires_grad_()
    torch.testing.assert_close(torch.autograd.grad(out, t, g)[0], torch.autograd.grad(_rotate_pairs(t2, freqs), t2, g)[0], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(R.fused_apply_rotary_pos_emb_cached(t, freqs.cos(), freqs.sin()), want, atol=1e-5, rtol=1e-5)
    lens = [rng.randint(1, 5) for _ in range(rng.randint(1, 4))]                      # packed sequences restart at position 0
    cu = torch.tensor([0] + list(itertools.accumulate(lens)), dtype=torch.int32)
    tt, fr = torch.randn(sum(lens), h, d), torch.randn(max(lens), 1, 1, r)
    pos = torch.cat([torch.arange(n) for n in lens])
    torch.testing.assert_close(R.fused_apply_rotary_pos_emb_thd(tt, cu, fr), _rotate_pairs(tt, fr[pos, 0]), atol=1e-5, rtol=1e-5)


def _all_alignments_nll(logp, label, T, U, blank):
    """-log of the summed probability of every monotone alignment: T blanks and U label emissions, the last step a blank."""
    terms = []
    for emit in itertools.combinations(range(T + U - 1), U):
        t = u = 0
        lp, ok = 0.0, True
        for step in range(T + U):
            if step in emit:
                lp += logp[t, u, label[u]].item()
                u += 1
            else:
                lp += logp[t, u, blank].item()
                t += 1
            if t == T and step != T + U - 1:
                ok = False
                break
        if ok:
            terms.append(lp)
    m = max(terms)
    return -(m + math.log(sum(math.exp(x - m) for x in terms)))


@pytest.mark.parametrize("seed", range(12))
def test_transducer_loss_reference_against_alignment_enumeration(seed):
    from apex_b200.contrib.transducer import TransducerLoss
    rng = random.Random(seed)
    torch.manual_seed(seed)
    B, V, Tm, Um = rng.randint(1, 3), rng.randint(2, 5), rng.randint(1, 4), rng.randint(0, 3)
    f_len = torch.tensor([rng.randint(1, Tm) for _ in range(B)], dtype=torch.int32)
    y_len = torch.tensor([rng.randint(0, Um) for _ in range(B)], dtype=torch.int32)
    f_len[rng.randrange(B)], y_len[rng.randrange(B)] = Tm, Um
    blank = rng.randrange(V)
    label = torch.tensor([[rng.choice([v for v in range(V) if v != blank]) for _ in range(Um)] for _ in range(B)], dtype=torch.int32).reshape(B, Um)
    x = torch.randn(B, Tm, Um + 1, V, requires_grad=True)
    loss = TransducerLoss()(x, label, f_len, y_len, blank)
    logp = torch.log_softmax(x.detach().double(), -1)
    for b in range(B):
        assert loss[b].item() == pytest.approx(_all_alignments_nll(logp[b], label[b].tolist(), int(f_len[b]), int(y_len[b]), blank), abs=1e-4)
    (g,) = torch.autograd.grad(loss.sum(), x)                                         # one entry by central differences
    idx, eps = (rng.randrange(B), 0, 0, rng.randrange(V)), 1e-3
    xp, xm = x.detach().clone(), x.detach().clone()
    xp[idx] += eps
    xm[idx] -= eps
    fd = (TransducerLoss()(xp, label, f_len, y_len, blank).sum() - TransducerLoss()(xm, label, f_len, y_len, blank).sum()) / (2 * eps)
    assert g[idx].item() == pytest.approx(fd.item(), abs=2e-3)


@pytest.mark.parametrize("seed", range(8))
def test_focal_loss_reference_against_the_closed_form(seed):
    """Element by element: loss = c_f (base + softplus(-x)), d loss / dx = c_f (c_b (base + softplus(-x)) - off_b), with the smoothing
    constants and the positive / negative coefficients of the reference kernel (focal_loss_cuda_kernel.cu:30-106), ignored rows (label -2)
    and padded classes contributing nothing."""
    from apex_b200.contrib.focal_loss import focal_loss
    rng = random.Random(seed)
    torch.manual_seed(seed)
    n, n_cls = rng.randint(1, 12), rng.randint(3, 9)
    real = rng.randint(2, n_cls)
    alpha, gamma, s = rng.choice([0.25, 0.5]), rng.choice([0.0, 1.5, 2.0]), rng.choice([0.0, 0.1])
    x = torch.randn(n, n_cls, requires_grad=True)
    tgt = torch.tensor([rng.choice([-2, -1] + list(range(real))) for _ in range(n)])
    npos = torch.tensor([float(rng.randint(1, 5))])
    loss = focal_loss(x, tgt, npos, real, alpha, gamma, s)
    (grad,) = torch.autograd.grad(loss, x)
    want, want_grad = 0.0, torch.zeros(n, n_cls, dtype=torch.float64)
    for i in range(n):
        if int(tgt[i]) == -2:
            continue
        for c in range(real):
            p = float(x[i, c].detach())
            sig, softplus = 1 / (1 + math.exp(-p)), math.log1p(math.exp(-abs(p))) + max(-p, 0.0)
            if int(tgt[i]) == c:
                base, off_b, c_f, c_b = (s - s / 2) * p, (1 - s + s / 2) - sig, alpha * (1 - sig) ** gamma, -gamma * sig
            else:
                base, off_b, c_f, c_b = (1 - s / 2) * p, s / 2 - sig, (1 - alpha) * sig ** gamma, gamma * (1 - sig)
            want += c_f * (base + softplus)
            want_grad[i, c] = c_f * (c_b * (base + softplus) - off_b)
    assert loss.item() == pytest.approx(want / float(npos), rel=1e-4, abs=1e-6)
    torch.testing.assert_close(grad.double(), want_grad / float(npos), atol=1e-5, rtol=1e-4)
