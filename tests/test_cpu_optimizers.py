"""CPU path: the PyTorch-reference fallbacks must match torch.optim (BASELINE.json config #1 plumbing)."""
import copy

import pytest
import torch

from apex_b200.normalization import FusedLayerNorm, FusedRMSNorm
from apex_b200.optimizers import FusedAdagrad, FusedAdam, FusedLAMB, FusedNovoGrad, FusedSGD


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g)) for s in [(37, 5), (278,), (4, 3, 3)]]


def _run(opt_a, opt_b, pa, pb, iters=7):
    g = torch.Generator().manual_seed(1)
    for _ in range(iters):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g)
            a.grad = gr.clone()
            b.grad = gr.clone()
        opt_a.step()
        opt_b.step()
    for a, b in zip(pa, pb):
        torch.testing.assert_close(a, b, atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("adam_w", [True, False])
def test_fused_adam_cpu_matches_torch(adam_w):
    pa, pb = _params(), _params()
    a = FusedAdam(pa, lr=1e-2, weight_decay=0.05, adam_w_mode=adam_w)
    b = (torch.optim.AdamW if adam_w else torch.optim.Adam)(pb, lr=1e-2, weight_decay=0.05)
    _run(a, b, pa, pb)


@pytest.mark.parametrize("nesterov", [False, True])
def test_fused_sgd_cpu_matches_torch(nesterov):
    pa, pb = _params(), _params()
    a = FusedSGD(pa, lr=1e-2, momentum=0.9, weight_decay=0.01, nesterov=nesterov)
    b = torch.optim.SGD(pb, lr=1e-2, momentum=0.9, weight_decay=0.01, nesterov=nesterov)
    _run(a, b, pa, pb)


def test_fused_adagrad_cpu_matches_torch():
    pa, pb = _params(), _params()
    _run(FusedAdagrad(pa, lr=1e-2, weight_decay=0.01), torch.optim.Adagrad(pb, lr=1e-2, weight_decay=0.01), pa, pb)


def test_mlp_trains_on_cpu():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(16, 32), FusedLayerNorm(32), torch.nn.ReLU(), torch.nn.Linear(32, 4), FusedRMSNorm(4))
    x, y = torch.randn(64, 16), torch.randn(64, 4)
    for Opt, kw in [(FusedAdam, dict(lr=1e-2)), (FusedLAMB, dict(lr=1e-2)), (FusedNovoGrad, dict(lr=1e-2))]:
        mm = copy.deepcopy(m)
        opt = Opt(mm.parameters(), **kw)
        first = last = None
        for _ in range(25):
            loss = ((mm(x) - y) ** 2).mean()
            first = first if first is not None else loss.item()
            opt.zero_grad()
            loss.backward()
            opt.step()
            last = loss.item()
        assert last < first, (Opt.__name__, first, last)


def test_state_dict_roundtrip_cpu():
    pa = _params()
    a = FusedAdam(pa, lr=1e-2)
    for p in pa:
        p.grad = torch.ones_like(p)
    a.step()
    sd = copy.deepcopy(a.state_dict())
    b = FusedAdam(_params(), lr=1e-2)
    b.load_state_dict(sd)
    assert b.state_dict()["param_groups"][0]["step"] == 1


def test_fused_lamb_cpu_matches_reference_lamb():
    """In-file LAMB oracle like the reference's RefLAMB (tests/L0/run_optimizers/test_lamb.py:11-100): global-norm clipping, Adam
    moments with bias correction, decoupled weight decay inside the update, per-tensor trust ratio."""
    import torch
    from apex_b200.optimizers import FusedLAMB
    torch.manual_seed(0)

    def ref_step(params, grads, state, lr, b1, b2, eps, wd, step, max_gn):
        gn = torch.sqrt(sum((g ** 2).sum() for g in grads))
        clip = max(float(gn) / max_gn, 1.0)
        for p, g, st in zip(params, grads, state):
            g = g / clip
            st["m"] = b1 * st["m"] + (1 - b1) * g
            st["v"] = b2 * st["v"] + (1 - b2) * g * g
            u = (st["m"] / (1 - b1 ** step)) / ((st["v"] / (1 - b2 ** step)).sqrt() + eps) + wd * p
            pn, un = p.norm(), u.norm()
            p -= lr * (float(pn / un) if (pn > 0 and un > 0) else 1.0) * u

    ps = [torch.nn.Parameter(torch.randn(37, 5)), torch.nn.Parameter(torch.randn(11))]
    qs = [p.detach().clone() for p in ps]
    st = [{"m": torch.zeros_like(q), "v": torch.zeros_like(q)} for q in qs]
    opt = FusedLAMB(ps, lr=1e-2, weight_decay=0.01, max_grad_norm=1.0)
    for step in range(1, 5):
        gs = [torch.randn_like(p) * 3 for p in ps]
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        opt.step()
        ref_step(qs, gs, st, 1e-2, 0.9, 0.999, 1e-6, 0.01, step, 1.0)
    for p, q in zip(ps, qs):
        torch.testing.assert_close(p.detach(), q, rtol=1e-5, atol=1e-6)


def test_novograd_and_mixed_precision_lamb_run_on_cpu():
    import torch
    from apex_b200.optimizers import FusedMixedPrecisionLamb, FusedNovoGrad
    torch.manual_seed(0)
    for cls in (FusedNovoGrad, FusedMixedPrecisionLamb):
        p = torch.nn.Parameter(torch.randn(37, 5))
        p0 = p.detach().clone()
        opt = cls([p], lr=1e-2)
        for _ in range(3):
            p.grad = torch.randn(37, 5)
            opt.step()
        assert torch.isfinite(p).all() and not torch.equal(p.detach(), p0)
        opt.load_state_dict(opt.state_dict())


def test_fused_novograd_cpu_matches_layerwise_oracle():
    """NovoGrad with a per-tensor second moment (same option set as tests/L0/run_optimizers/test_fused_novograd.py:130-155:
    betas (0.95, 0), no bias correction, norm inside the moment, L2 norm, no zero init)."""
    import torch
    from apex_b200.optimizers import FusedNovoGrad
    torch.manual_seed(0)
    shapes = [(64, 9), (31,), (1,)]
    ps = [torch.nn.Parameter(torch.rand(s)) for s in shapes]
    qs = [p.detach().clone() for p in ps]
    m = [torch.zeros_like(q) for q in qs]
    v = [torch.zeros(()) for _ in qs]
    opt = FusedNovoGrad(ps, lr=1e-3, betas=(0.95, 0), eps=1e-8, weight_decay=0, grad_averaging=False, bias_correction=False,
                        reg_inside_moment=True, norm_type=2, init_zero=False)
    for _ in range(5):
        gs = [torch.rand_like(q) for q in qs]
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        opt.step()
        for i, (q, g) in enumerate(zip(qs, gs)):
            n2 = (g * g).sum()
            v[i] = n2 if float(v[i]) == 0 else 0.0 * v[i] + 1.0 * n2
            m[i] = 0.95 * m[i] + g / (v[i].sqrt() + 1e-8)
            q -= 1e-3 * m[i]
    for p, q in zip(ps, qs):
        torch.testing.assert_close(p.detach(), q, rtol=1e-5, atol=1e-7)


def _resume_factories():
    from apex_b200.contrib.optimizers import DistributedFusedAdam, DistributedFusedLAMB
    from apex_b200.contrib.optimizers import FusedAdam as LegacyAdam
    from apex_b200.contrib.optimizers import FusedLAMB as LegacyLAMB
    from apex_b200.contrib.optimizers import FusedSGD as LegacySGD
    from apex_b200.optimizers import FusedAdagrad, FusedAdam, FusedLAMB, FusedMixedPrecisionLamb, FusedNovoGrad, FusedSGD

    return {
        "FusedAdam": lambda ps: FusedAdam(ps, lr=1e-2, weight_decay=0.01),
        "FusedLAMB": lambda ps: FusedLAMB(ps, lr=1e-2, weight_decay=0.01),
        "FusedMixedPrecisionLamb": lambda ps: FusedMixedPrecisionLamb(ps, lr=1e-2, weight_decay=0.01),
        "FusedSGD": lambda ps: FusedSGD(ps, lr=1e-2, momentum=0.9, dampening=0.1, weight_decay=0.01),
        "FusedNovoGrad": lambda ps: FusedNovoGrad(ps, lr=1e-2, weight_decay=0.01),
        "FusedAdagrad": lambda ps: FusedAdagrad(ps, lr=1e-2, weight_decay=0.01),
        "DistributedFusedAdam": lambda ps: DistributedFusedAdam(ps, lr=1e-2, weight_decay=0.01, device="cpu"),
        "DistributedFusedLAMB": lambda ps: DistributedFusedLAMB(ps, lr=1e-2, weight_decay=0.01, device="cpu"),
        "contrib.FusedAdam": lambda ps: LegacyAdam(ps, lr=1e-2, weight_decay=0.01),
        "contrib.FusedLAMB": lambda ps: LegacyLAMB(ps, lr=1e-2, weight_decay=0.01),
        "contrib.FusedSGD": lambda ps: LegacySGD(ps, lr=1e-2, momentum=0.9),
    }


@pytest.mark.parametrize("name", ["FusedAdam", "FusedLAMB", "FusedMixedPrecisionLamb", "FusedSGD", "FusedNovoGrad", "FusedAdagrad",
                                  "DistributedFusedAdam", "DistributedFusedLAMB", "contrib.FusedAdam", "contrib.FusedLAMB", "contrib.FusedSGD"])
def test_checkpoint_resume_is_exact(name):
    """6 uninterrupted steps == 3 steps, state_dict -> fresh optimizer on fresh parameters, 3 more steps (moments, step counters,
    per-tensor norms and sharded state must all survive the round trip)."""
    import copy
    import warnings

    make = _resume_factories()[name]

    def fresh():
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.randn(70)), torch.nn.Parameter(torch.randn(3, 5))]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return ps, make(ps)

    def run(ps, opt, its):
        for it in its:
            g = torch.Generator().manual_seed(it)
            opt.zero_grad()
            for p in ps:
                p.grad = torch.randn(p.shape, generator=g) * (it + 1)
            opt.step()

    ps, opt = fresh()
    run(ps, opt, range(6))
    ps_a, opt_a = fresh()
    run(ps_a, opt_a, range(3))
    sd = copy.deepcopy(opt_a.state_dict())
    ps_b, opt_b = fresh()
    with torch.no_grad():
        for dst, src in zip(ps_b, ps_a):
            dst.copy_(src)
    opt_b.load_state_dict(sd)
    run(ps_b, opt_b, range(3, 6))
    for got, want in zip(ps_b, ps):
        torch.testing.assert_close(got, want, rtol=0, atol=0)


@pytest.mark.parametrize("momentum,dampening,nesterov", [(0.9, 0.0, False), (0.9, 0.3, False), (0.9, 0.0, True), (0.0, 0.0, False)])
def test_fused_sgd_parameters_whose_first_gradient_arrives_late(momentum, dampening, nesterov):
    """torch.optim.SGD semantics when a parameter has no gradient on some steps: its momentum buffer starts as the first gradient IT sees,
    and the buffers of the other parameters are not disturbed (the launch-wide ``first_run`` flag must not leak across parameters)."""
    for skip in (0, 1, 2):
        torch.manual_seed(1)
        pa = [torch.nn.Parameter(torch.randn(11)), torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(2, 2, 2))]
        pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
        kw = dict(lr=0.1, momentum=momentum, dampening=dampening, nesterov=nesterov, weight_decay=0.1)
        a = FusedSGD([{"params": pa[:1]}, {"params": pa[1:], "lr": 0.05}], **kw)
        b = torch.optim.SGD([{"params": pb[:1]}, {"params": pb[1:], "lr": 0.05}], **kw)
        for it in range(5):
            g = torch.Generator().manual_seed(it)
            for i, (x, y) in enumerate(zip(pa, pb)):
                grad = torch.randn(x.shape, generator=g)
                x.grad, y.grad = (None, None) if (i == skip and it % 2 == 0) else (grad.clone(), grad.clone())
            a.step()
            b.step()
        for x, y in zip(pa, pb):
            torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["adam_bf16_params", "adam_capturable_master_bf16", "mixed_precision_lamb_bf16"])
def test_checkpoint_resume_is_exact_with_16_bit_parameters(name):
    """fp32 moments and fp32 master weights of 16-bit parameters must come back at full precision: torch's load_state_dict casts state to the
    parameter dtype, and the reference keeps master weights out of the checkpoint — both made a resumed run drift by bf16 ulps."""
    import copy

    from apex_b200.optimizers import FusedMixedPrecisionLamb

    def fresh():
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.randn(70).bfloat16()), torch.nn.Parameter(torch.randn(3, 5).bfloat16())]
        if name == "adam_bf16_params":
            return ps, FusedAdam(ps, lr=1e-2, weight_decay=0.01)
        if name == "adam_capturable_master_bf16":
            return ps, FusedAdam(ps, lr=1e-2, weight_decay=0.01, capturable=True, master_weights=True)
        return ps, FusedMixedPrecisionLamb(ps, lr=1e-2, weight_decay=0.01, reduced_precision_dtype=torch.bfloat16)

    def run(ps, opt, its):
        for it in its:
            g = torch.Generator().manual_seed(it)
            opt.zero_grad()
            for p in ps:
                p.grad = (torch.randn(p.shape, generator=g) * (it + 1)).bfloat16()
            opt.step()

    ps, opt = fresh()
    run(ps, opt, range(6))
    pa, oa = fresh()
    run(pa, oa, range(3))
    sd = copy.deepcopy(oa.state_dict())
    pb, ob = fresh()
    with torch.no_grad():
        for dst, src in zip(pb, pa):
            dst.copy_(src)
    ob.load_state_dict(sd)
    run(pb, ob, range(3, 6))
    for got, want in zip(pb, ps):
        torch.testing.assert_close(got, want, rtol=0, atol=0)


@pytest.mark.parametrize("family", ["adamw", "adamw_capturable", "sgd", "adagrad"])
def test_checkpoints_of_the_torch_optimizers_continue_in_the_fused_ones(family):
    """Migration path: train with torch.optim, load that optimizer's state_dict into the fused counterpart, keep training: identical
    trajectory (missing apex-only hyper-parameters come from the defaults, per-parameter step counters become the per-group counter)."""
    import copy

    torch_make, fused_make = {
        "adamw": (lambda p: torch.optim.AdamW(p, lr=1e-2, weight_decay=0.01), lambda p: FusedAdam(p, lr=1e-2, weight_decay=0.01)),
        "adamw_capturable": (lambda p: torch.optim.AdamW(p, lr=1e-2, weight_decay=0.01), lambda p: FusedAdam(p, lr=1e-2, weight_decay=0.01, capturable=True)),
        "sgd": (lambda p: torch.optim.SGD(p, lr=1e-2, momentum=0.9, weight_decay=0.01), lambda p: FusedSGD(p, lr=1e-2, momentum=0.9, weight_decay=0.01)),
        "adagrad": (lambda p: torch.optim.Adagrad(p, lr=1e-2), lambda p: FusedAdagrad(p, lr=1e-2)),
    }[family]

    def grads(ps, it):
        g = torch.Generator().manual_seed(it)
        for p in ps:
            p.grad = torch.randn(p.shape, generator=g)

    torch.manual_seed(0)
    pa = [torch.nn.Parameter(torch.randn(9)), torch.nn.Parameter(torch.randn(3, 4))]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ta, tb = torch_make(pa), torch_make(pb)
    for it in range(3):
        grads(pa, it)
        ta.step()
        grads(pb, it)
        tb.step()
    fused = fused_make(pa)
    fused.load_state_dict(copy.deepcopy(ta.state_dict()))
    for it in range(3, 6):
        grads(pa, it)
        fused.step()
        grads(pb, it)
        tb.step()
    for x, y in zip(pa, pb):
        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-6)
