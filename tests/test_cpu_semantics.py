"""Behavioural sweeps against PyTorch's own modules / optimizers on the device-independent (PyTorch) paths: option combinations a single
example test does not reach. The CUDA kernels are checked against the same oracles in tests/test_gpu_*.py."""
import itertools
import warnings

import pytest
import torch
import torch.nn.functional as F


def test_layer_norm_option_sweep():
    from apex_b200.normalization import FusedLayerNorm
    torch.manual_seed(0)
    for shape, ns, affine, me in itertools.product([(4, 8), (2, 3, 8), (2, 3, 4, 8)], [1, 2], [True, False], [False, True]):
        nshape = shape[-ns:]
        a, b = FusedLayerNorm(nshape, elementwise_affine=affine, memory_efficient=me), torch.nn.LayerNorm(nshape, elementwise_affine=affine)
        if affine:
            with torch.no_grad():
                a.weight.normal_()
                a.bias.normal_()
                b.weight.copy_(a.weight)
                b.bias.copy_(a.bias)
        x = torch.randn(shape)
        xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
        ya, yb = a(xa), b(xb)
        dy = torch.randn_like(ya)
        ya.backward(dy)
        yb.backward(dy)
        torch.testing.assert_close(ya, yb, atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(xa.grad, xb.grad, atol=1e-4, rtol=1e-4)
        if affine:
            torch.testing.assert_close(a.weight.grad, b.weight.grad, atol=1e-4, rtol=1e-4)
            torch.testing.assert_close(a.bias.grad, b.bias.grad, atol=1e-4, rtol=1e-4)
    with pytest.raises(RuntimeError):
        FusedLayerNorm(8)(torch.randn(2, 7))


def test_clip_grad_norm_types_and_nonfinite():
    from apex_b200.contrib.clip_grad import clip_grad_norm_
    torch.manual_seed(0)
    for norm_type, max_norm in itertools.product([2.0, float("inf"), 1.0, 3.0], [0.5, 100.0]):
        ps = [torch.nn.Parameter(torch.randn(5)), torch.nn.Parameter(torch.randn(3, 2))]
        pr = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        for p, r in zip(ps, pr):
            p.grad = torch.randn_like(p)
            r.grad = p.grad.clone()
        torch.testing.assert_close(clip_grad_norm_(ps, max_norm, norm_type=norm_type), torch.nn.utils.clip_grad_norm_(pr, max_norm, norm_type=norm_type))
        for p, r in zip(ps, pr):
            torch.testing.assert_close(p.grad, r.grad)
    p = torch.nn.Parameter(torch.randn(5))
    p.grad = torch.full((5,), float("nan"))
    with pytest.raises(RuntimeError):
        clip_grad_norm_([p], 1.0, error_if_nonfinite=True)


def test_xentropy_smoothing_and_padding_sweep():
    from apex_b200.contrib.xentropy import SoftmaxCrossEntropyLoss
    torch.manual_seed(0)
    for smoothing, pad in itertools.product([0.0, 0.1], [None, 0, 3]):
        logits, labels = torch.randn(9, 7), torch.randint(0, 7, (9,))
        if pad is not None:
            labels[2] = pad
        la, lb = logits.clone().requires_grad_(), logits.clone().requires_grad_()
        got = SoftmaxCrossEntropyLoss.apply(la, labels, smoothing, -1 if pad is None else pad, True)
        want = F.cross_entropy(lb, labels, label_smoothing=smoothing, reduction="none", ignore_index=-100 if pad is None else pad)
        got.sum().backward()
        want.sum().backward()
        torch.testing.assert_close(got, want, atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(la.grad, lb.grad, atol=1e-5, rtol=1e-5)


def _rotate_half(t):
    d = t.shape[-1] // 2
    return torch.cat((-t[..., d:], t[..., :d]), -1)


def _rope(t, freqs):
    r = freqs.shape[-1]
    return torch.cat((t[..., :r] * freqs.cos() + _rotate_half(t[..., :r]) * freqs.sin(), t[..., r:]), -1)


def test_rope_variants():
    from apex_b200.transformer.functional import fused_rope as R
    torch.manual_seed(0)
    s, b, h, d = 6, 2, 3, 8
    for r in (8, 4):   # full and partial rotary dimension
        t, f = torch.randn(s, b, h, d, requires_grad=True), torch.randn(s, 1, 1, r)
        out, want = R.fused_apply_rotary_pos_emb(t, f), _rope(t, f)
        g = torch.randn_like(out)
        torch.testing.assert_close(out, want, atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(torch.autograd.grad(out, t, g)[0], torch.autograd.grad(want, t, g)[0], atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(R.fused_apply_rotary_pos_emb_cached(t, f.cos(), f.sin()), want, atol=1e-5, rtol=1e-5)
        tr = R.fused_apply_rotary_pos_emb(t, f, transpose_output_memory=True)
        torch.testing.assert_close(tr, want, atol=1e-5, rtol=1e-5)
        assert tr.transpose(0, 1).is_contiguous()
    cu = torch.tensor([0, 3, 7, 12], dtype=torch.int32)
    tt, f = torch.randn(12, h, d, requires_grad=True), torch.randn(6, 1, 1, d)
    out = R.fused_apply_rotary_pos_emb_thd(tt, cu, f)
    want = torch.cat([_rope(tt[cu[i]:cu[i + 1]].unsqueeze(1), f[:cu[i + 1] - cu[i]]).squeeze(1) for i in range(3)])
    g = torch.randn_like(out)
    torch.testing.assert_close(out, want, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(torch.autograd.grad(out, tt, g)[0], torch.autograd.grad(want, tt, g)[0], atol=1e-5, rtol=1e-5)
    bb, ih, iw = 2, 3, 4
    t2 = torch.randn(bb, ih * iw, h, d)
    ch, sh, cw, sw = torch.randn(1, ih + 1, 1, d // 2), torch.randn(1, ih + 1, 1, d // 2), torch.randn(1, iw + 2, 1, d // 2), torch.randn(1, iw + 2, 1, d // 2)
    x = t2.view(bb, ih, iw, h, d)
    xh, xw = x[..., :d // 2], x[..., d // 2:]
    want = torch.cat((xh * ch[:, :ih].unsqueeze(2) + _rotate_half(xh) * sh[:, :ih].unsqueeze(2),
                      xw * cw[:, :iw].unsqueeze(1) + _rotate_half(xw) * sw[:, :iw].unsqueeze(1)), -1).view(bb, ih * iw, h, d)
    torch.testing.assert_close(R.fused_apply_rotary_pos_emb_2d(t2, ih, iw, ch, sh, cw, sw), want, atol=1e-5, rtol=1e-5)


def _three_params(dtype=torch.float32):
    torch.manual_seed(1)
    return [torch.nn.Parameter(torch.randn(1100).to(dtype)), torch.nn.Parameter(torch.randn(40, 30).to(dtype)), torch.nn.Parameter(torch.randn(7).to(dtype))]


def _two_groups(ps):
    return [{"params": ps[:2], "lr": 1e-2}, {"params": ps[2:], "lr": 3e-3, "weight_decay": 0.05, "betas": (0.8, 0.9)}]


@pytest.mark.parametrize("adam_w", [True, False])
def test_adam_family_param_groups_match_torch(adam_w):
    """FusedAdam and DistributedFusedAdam (several bucket / buffer configurations, deferred clipping) with per-group lr / weight decay /
    betas against torch.optim.Adam(W)."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    from apex_b200.optimizers import FusedAdam
    torch_cls = torch.optim.AdamW if adam_w else torch.optim.Adam
    makers = [lambda ps: FusedAdam(_two_groups(ps), lr=1e-3, weight_decay=0.1, adam_w_mode=adam_w)]
    for kw in (dict(), dict(bucket_cap_mb=0.001), dict(contiguous_grad_buffer=False, contiguous_param_buffer=False), dict(overlap_grad_sync=False)):
        makers.append(lambda ps, kw=kw: DistributedFusedAdam(_two_groups(ps), lr=1e-3, weight_decay=0.1, adam_w_mode=adam_w, device="cpu", **kw))
    for make in makers:
        pa, pb = _three_params(), _three_params()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = make(pa)
        b = torch_cls(_two_groups(pb), lr=1e-3, weight_decay=0.1)
        for it in range(4):
            g = torch.Generator().manual_seed(it)
            a.zero_grad()
            b.zero_grad()
            for x, y in zip(pa, pb):
                grad = torch.randn(x.shape, generator=g)
                x.grad = grad.clone() if x.grad is None else x.grad.copy_(grad)
                y.grad = grad.clone()
            if it == 2 and hasattr(a, "clip_grad_norm"):
                torch.testing.assert_close(a.clip_grad_norm(0.5).reshape(()), torch.nn.utils.clip_grad_norm_(pb, 0.5).reshape(()), atol=1e-4, rtol=1e-4)
            a.step()
            b.step()
        for x, y in zip(pa, pb):
            torch.testing.assert_close(x, y, atol=2e-6, rtol=1e-5)


def test_adagrad_matches_torch():
    from apex_b200.optimizers import FusedAdagrad
    for wd in (0.0, 0.1):
        pa, pb = _three_params(), _three_params()
        a, b = FusedAdagrad(pa, lr=0.1, weight_decay=wd), torch.optim.Adagrad(pb, lr=0.1, weight_decay=wd)
        for it in range(4):
            g = torch.Generator().manual_seed(it)
            for x, y in zip(pa, pb):
                x.grad = torch.randn(x.shape, generator=g)
                y.grad = x.grad.clone()
            a.step()
            b.step()
        for x, y in zip(pa, pb):
            torch.testing.assert_close(x, y, atol=1e-6, rtol=1e-5)


def test_optimizer_argument_validation():
    from apex_b200.optimizers import FusedAdagrad, FusedAdam, FusedLAMB, FusedNovoGrad, FusedSGD
    p = [torch.nn.Parameter(torch.randn(4, 4))]
    for cls in (FusedAdam, FusedLAMB, FusedNovoGrad):
        with pytest.raises(RuntimeError, match="AMSGrad"):
            cls(p, amsgrad=True)
    with pytest.raises(ValueError):
        FusedSGD(p, lr=-1)
    with pytest.raises(ValueError):
        FusedSGD(p, lr=0.1, nesterov=True)
    for cls, kw in ((FusedAdam, {}), (FusedLAMB, {}), (FusedSGD, {"lr": 0.1}), (FusedAdagrad, {})):
        q = [torch.nn.Parameter(torch.randn(4, 4))]
        opt = cls(q, **kw)
        q[0].grad = torch.randn(4, 4).to_sparse()
        with pytest.raises(RuntimeError, match="sparse"):
            opt.step()


@pytest.mark.parametrize("name", ["FusedAdam", "FusedLAMB", "FusedSGD", "FusedNovoGrad", "FusedAdagrad", "FusedMixedPrecisionLamb",
                                  "DistributedFusedAdam", "DistributedFusedLAMB"])
def test_grad_scaler_skips_the_step_on_overflow(name):
    """torch.amp.GradScaler drives every optimizer: an inf gradient must leave the parameters untouched and halve the scale, finite steps
    must update (reference tests/L0/run_optimizers/test_adam.py trains a small net under a GradScaler)."""
    from apex_b200.contrib import optimizers as CO
    from apex_b200 import optimizers as O
    kw = {"FusedSGD": dict(lr=1e-2, momentum=0.9)}.get(name, dict(lr=1e-2))
    if name.startswith("Distributed"):
        kw["device"] = "cpu"
    torch.manual_seed(0)
    model = torch.nn.Linear(8, 4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt = getattr(CO if name.startswith("Distributed") else O, name)(model.parameters(), **kw)
    scaler = torch.amp.GradScaler("cpu", init_scale=1024.0, growth_interval=2)
    for it in range(5):
        opt.zero_grad()
        scaler.scale(model(torch.randn(16, 8)).square().mean()).backward()
        before = [p.detach().clone() for p in model.parameters()]
        if it == 2:
            for p in model.parameters():
                p.grad[0] = float("inf")
        scaler.step(opt)
        scaler.update()
        changed = any(not torch.equal(a, b.detach()) for a, b in zip(before, model.parameters()))
        assert changed == (it != 2), (name, it)
        if it == 2:
            assert scaler.get_scale() == 1024.0    # grew to 2048 after two clean steps, halved by the overflow
    assert all(bool(torch.isfinite(p).all()) for p in model.parameters())


def test_lamb_option_sweep_against_an_independent_oracle():
    """FusedLAMB over adam_w_mode x use_nvlamb x grad_averaging x bias_correction x (clipping on / off) with two param groups: the gradient
    norm is global over the groups, the trust ratio applies only where weight decay != 0 unless use_nvlamb (reference fused_lamb.py +
    csrc/multi_tensor_lamb.cu semantics, written out independently here)."""
    from apex_b200.optimizers import FusedLAMB

    def oracle(groups, state, step, max_gn, adam_w, nvlamb, grad_avg, bias_corr):
        gn = torch.sqrt(sum((g ** 2).sum() for grp in groups for g in grp["grads"]))
        clip = max(float(gn) / max_gn, 1.0) if max_gn > 0 else 1.0
        for grp in groups:
            b1, b2 = grp["betas"]
            b3 = 1 - b1 if grad_avg else 1.0
            for p, g in zip(grp["params"], grp["grads"]):
                st = state[id(p)]
                g = g / clip
                if not adam_w:
                    g = g + grp["wd"] * p
                st["m"] = b1 * st["m"] + b3 * g
                st["v"] = b2 * st["v"] + (1 - b2) * g * g
                c1, c2 = (1 - b1 ** step, 1 - b2 ** step) if bias_corr else (1.0, 1.0)
                u = (st["m"] / c1) / ((st["v"] / c2).sqrt() + grp["eps"])
                if adam_w:
                    u = u + grp["wd"] * p
                ratio = 1.0
                if nvlamb or grp["wd"] != 0:
                    pn, un = p.norm(), u.norm()
                    ratio = float(pn / un) if (pn > 0 and un > 0) else 1.0
                p -= grp["lr"] * ratio * u

    for adam_w, nvlamb, grad_avg, bias_corr, max_gn in itertools.product([True, False], [False, True], [True, False], [True, False], [1.0, 0.0]):
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.randn(37, 5)), torch.nn.Parameter(torch.randn(11)), torch.nn.Parameter(torch.randn(4, 4))]
        qs = [p.detach().clone() for p in ps]
        opt = FusedLAMB([{"params": ps[:2], "weight_decay": 0.01}, {"params": ps[2:], "weight_decay": 0.0, "lr": 3e-3, "betas": (0.8, 0.95)}], lr=1e-2,
                        eps=1e-6, adam_w_mode=adam_w, use_nvlamb=nvlamb, grad_averaging=grad_avg, bias_correction=bias_corr, max_grad_norm=max_gn)
        state = {id(q): {"m": torch.zeros_like(q), "v": torch.zeros_like(q)} for q in qs}
        for step in range(1, 4):
            gs = [torch.randn_like(p) * 3 for p in ps]
            for p, g in zip(ps, gs):
                p.grad = g.clone()
            opt.step()
            oracle([{"params": qs[:2], "grads": gs[:2], "wd": 0.01, "lr": 1e-2, "betas": (0.9, 0.999), "eps": 1e-6},
                    {"params": qs[2:], "grads": gs[2:], "wd": 0.0, "lr": 3e-3, "betas": (0.8, 0.95), "eps": 1e-6}], state, step, max_gn, adam_w, nvlamb,
                   grad_avg, bias_corr)
        for p, q in zip(ps, qs):
            torch.testing.assert_close(p.detach(), q, rtol=1e-4, atol=2e-5)
