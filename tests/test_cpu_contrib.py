"""CPU checks of the contrib modules: import surface, small numerics against plain PyTorch (the CUDA kernels have GPU twins)."""
import importlib

import pytest
import torch
import torch.nn as nn


@pytest.mark.parametrize("mod", [
    "apex_b200.contrib.optimizers", "apex_b200.contrib.xentropy", "apex_b200.contrib.layer_norm", "apex_b200.contrib.group_norm",
    "apex_b200.contrib.clip_grad", "apex_b200.contrib.focal_loss", "apex_b200.contrib.index_mul_2d", "apex_b200.contrib.transducer",
    "apex_b200.contrib.multihead_attn", "apex_b200.contrib.fmha", "apex_b200.contrib.conv_bias_relu", "apex_b200.contrib.bottleneck",
    "apex_b200.contrib.groupbn", "apex_b200.contrib.cudnn_gbn", "apex_b200.contrib.sparsity", "apex_b200.contrib.peer_memory",
    "apex_b200.contrib.nccl_p2p", "apex_b200.contrib.nccl_allocator", "apex_b200.contrib.gpu_direct_storage", "apex_b200.contrib.openfold",
    "apex_b200.contrib.torchsched", "apex_b200.parallel", "apex_b200.transformer.functional", "apex_b200.fused_dense", "apex_b200.mlp",
    "apex_b200.normalization", "apex_b200.optimizers", "apex_b200.multi_tensor_apply", "apex_b200.utils.flatten"])
def test_imports(mod):
    importlib.import_module(mod)


def test_transducer_loss_oracle_matches_bruteforce():
    from apex_b200.contrib.transducer import TransducerLoss
    torch.manual_seed(0)
    B, T, U, V = 2, 4, 3, 5
    x = torch.randn(B, T, U, V, requires_grad=True)
    label = torch.randint(1, V, (B, U - 1))
    f_len, y_len = torch.tensor([4, 3]), torch.tensor([2, 1])
    loss = TransducerLoss()(x, label, f_len, y_len, 0)
    lp = torch.log_softmax(x.detach(), -1)

    def brute(b):
        Tb, Ub = int(f_len[b]), int(y_len[b]) + 1

        def rec(t, u):  # log-prob of finishing from (t, u)
            if t == Tb - 1 and u == Ub - 1:
                return lp[b, t, u, 0]
            opts = []
            if t < Tb - 1:
                opts.append(lp[b, t, u, 0] + rec(t + 1, u))
            if u < Ub - 1:
                opts.append(lp[b, t, u, label[b, u]] + rec(t, u + 1))
            return torch.logsumexp(torch.stack(opts), 0)
        return -rec(0, 0)

    torch.testing.assert_close(loss.detach(), torch.stack([brute(0), brute(1)]), atol=1e-5, rtol=1e-5)
    loss.sum().backward()
    assert torch.isfinite(x.grad).all()


def test_transducer_joint_cpu():
    from apex_b200.contrib.transducer import TransducerJoint
    f, g = torch.randn(2, 5, 8), torch.randn(2, 3, 8)
    f_len, g_len = torch.tensor([5, 3]), torch.tensor([3, 2])
    out = TransducerJoint(relu=True)(f, g, f_len, g_len)
    assert out.shape == (2, 5, 3, 8)
    torch.testing.assert_close(out[0], torch.relu(f[0, :, None] + g[0, None]))
    assert out[1, 3:].abs().sum() == 0 and out[1, :, 2:].abs().sum() == 0
    bo = torch.cumsum(f_len * g_len, 0)
    packed = TransducerJoint(pack_output=True)(f, g, f_len, g_len, batch_offset=bo, packed_batch=int(bo[-1]))
    assert packed.shape == (21, 8)


def test_permutation_search_improves_and_is_a_permutation():
    from apex_b200.contrib.sparsity import permutation_search as P
    assert len(P.generate_all_unique_combinations(8)) == 35
    torch.manual_seed(0)
    m = torch.randn(64, 32)
    base = float(P.sum_after_2_to_4(m))
    for fn in (lambda: P.Channel_Swap(m), lambda: P.Exhaustive_Search(m, 8), lambda: P.Random_Search(m, 20)):
        out, _, perm = fn()
        assert sorted(perm) == list(range(32))
        assert float(P.sum_after_2_to_4(out)) >= base
        torch.testing.assert_close(out, m[:, perm])
    out, _, _ = P.Exhaustive_Search(m, 8)
    assert float(P.sum_after_2_to_4(out)) > base * 1.005


def test_asp_with_permutation_preserves_function():
    from apex_b200.contrib.sparsity import ASP
    ASP.reset()
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 16))
    x = torch.randn(4, 32)
    y0 = model(x).detach()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    from apex_b200.contrib.sparsity.permutation_lib import Permutation
    Permutation.search_options = {"strategy": "exhaustive", "stripe_group_size": 8, "escape_attempts": 1}
    rep = Permutation.permute_model(model)
    assert len(rep) >= 1 and all(a > b for _, b, a in rep)
    torch.testing.assert_close(model(x), y0, atol=1e-5, rtol=1e-5)
    ASP.init_model_for_pruning(model, "m4n2_1d", verbosity=0, whitelist=(nn.Linear,), allow_recompute_mask=False)
    ASP.init_optimizer_for_pruning(opt)
    ASP.compute_sparse_masks()
    w = model[2].weight
    assert (w.view(-1, 4) != 0).sum(1).max() <= 2
    ASP.reset()


def test_syncbn_functional_cpu():
    from apex_b200.parallel import syncbn_ops as S
    x = torch.randn(4, 6, 5, 5)
    mean, var = S.welford_mean_var(x)
    m2, var_u, istd = S.welford_parallel(torch.stack([mean, mean]), torch.stack([var, var]), torch.tensor([100, 100]), 1e-5)
    torch.testing.assert_close(m2, mean)
    w, b = torch.randn(6), torch.randn(6)
    y = S.batchnorm_forward(x, mean, istd, w, b)
    torch.testing.assert_close(y, torch.nn.functional.batch_norm(x, None, None, w, b, True, 0.0, 1e-5), atol=1e-5, rtol=1e-5)


def test_legacy_contrib_optimizers_cpu():
    from apex_b200.contrib.optimizers import FP16_Optimizer, FusedAdam
    from apex_b200.optimizers import FusedAdam as FA
    p = torch.nn.Parameter(torch.ones(10))
    o = FusedAdam([p], lr=0.1)
    o.step(grads=[torch.full((10,), 4.0)], scale=2.0)
    torch.testing.assert_close(p.detach(), torch.full((10,), 0.9))
    m = torch.nn.Linear(4, 4).half()
    opt = FP16_Optimizer(FA(m.parameters(), lr=1e-2), static_loss_scale=8.0, verbose=False)
    w0 = m.weight.detach().clone()
    opt.backward(m(torch.randn(2, 4).half()).float().sum())
    opt.step()
    assert not torch.equal(w0, m.weight)
    sd = opt.state_dict()
    opt.load_state_dict(sd)


def test_legacy_contrib_fused_sgd_explicit_grads():
    """apex.contrib.optimizers.FusedSGD.step(grads=, output_params=, scale=): fp32 masters updated from scaled half gradients, half model
    copies written by the same call, fp32 model weights in the same group take the 3-list launch."""
    from apex_b200.contrib.optimizers import FusedSGD
    torch.manual_seed(0)
    masters = [nn.Parameter(torch.randn(33)), nn.Parameter(torch.randn(7, 5))]
    model = [masters[0].detach().half(), masters[1].detach().clone()]
    refs = [nn.Parameter(m.detach().clone()) for m in masters]
    o = FusedSGD(masters, lr=0.1, momentum=0.9, weight_decay=0.01)
    r = torch.optim.SGD(refs, lr=0.1, momentum=0.9, weight_decay=0.01)
    for _ in range(3):
        gs = [torch.randn(33), torch.randn(7, 5)]
        for q, g in zip(refs, gs):
            q.grad = g.clone()
        r.step()
        o.step(grads=[(gs[0] * 16).half(), gs[1] * 16], output_params=model, scale=16.0)
    torch.testing.assert_close(masters[0].detach(), refs[0].detach(), atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(masters[1].detach(), refs[1].detach(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(model[0], masters[0].detach().half())
    with pytest.raises(RuntimeError):
        o.step(grads=[gs[0], gs[1]])
    for m in masters:       # without explicit arguments it is the modern FusedSGD
        m.grad = torch.ones_like(m)
    o.step()


def test_fused_adam_swa_tracks_adam_and_averages():
    from apex_b200.contrib.openfold import FusedAdamSWA
    torch.manual_seed(0)
    ps = [nn.Parameter(torch.randn(10)) for _ in range(2)]
    cs = [nn.Parameter(p.detach().clone()) for p in ps]
    ss = [p.detach().clone() for p in ps]
    o = FusedAdamSWA(ps, cs, ss, swa_decay_rate=0.9, lr=1e-2)
    qs = [nn.Parameter(p.detach().clone()) for p in ps]
    ro = torch.optim.Adam(qs, lr=1e-2)
    for _ in range(3):
        for c, q in zip(cs, qs):
            g = torch.randn(10)
            c.grad, q.grad = g.clone(), g * 0.5
        o.step(grad_clip_scale=0.5)
        ro.step()
    for p, q, c in zip(ps, qs, cs):
        torch.testing.assert_close(p, q, atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(c.detach(), p.detach())


def test_asp_masks_survive_checkpoint_and_optimizer_steps():
    """Masks are module buffers: they travel with model.state_dict() and keep the weights 2:4 sparse through optimizer steps
    (reference apex/contrib/sparsity/test/checkpointing_test_part{1,2}.py)."""
    from apex_b200.contrib.sparsity import ASP

    def make():
        torch.manual_seed(0)
        return nn.Sequential(nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 16))

    ASP.reset()
    model = make()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    ASP.init_model_for_pruning(model, "m4n2_1d", verbosity=0, whitelist=(nn.Linear,), allow_recompute_mask=True)
    ASP.init_optimizer_for_pruning(opt)
    ASP.compute_sparse_masks()
    for _ in range(3):
        opt.zero_grad()
        model(torch.randn(8, 32)).pow(2).mean().backward()
        opt.step()
    w = model[0].weight
    assert int((w.view(-1, 4) != 0).sum(1).max()) <= 2, "optimizer step must re-apply the masks"
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    assert any("mma_mask" in k for k in sd)
    ASP.reset()
    model2 = make()
    opt2 = torch.optim.SGD(model2.parameters(), lr=0.1)
    ASP.init_model_for_pruning(model2, "m4n2_1d", verbosity=0, whitelist=(nn.Linear,), allow_recompute_mask=True)
    ASP.init_optimizer_for_pruning(opt2)
    model2.load_state_dict(sd)
    for k, v in model2.state_dict().items():
        torch.testing.assert_close(v, sd[k])
    opt2.zero_grad()
    model2(torch.randn(8, 32)).pow(2).mean().backward()
    opt2.step()
    assert int((model2[0].weight.view(-1, 4) != 0).sum(1).max()) <= 2
    ASP.restore_pruned_weights()
    ASP.reset()


def test_fused_adam_cuda_entry_points_reversible_step_and_checks():
    """reversible_adam followed by maybe_adam_undo (flag set) restores p, m, v; the strided finite check samples every k-th element
    (reference apex/contrib/csrc/optimizers/fused_adam_cuda.cpp:92-104)."""
    from apex_b200.contrib.optimizers import fused_adam_cuda as F
    torch.manual_seed(0)
    for mode in (0, 1):
        p, m, v, g = torch.randn(100), torch.rand(100) * 0.1, torch.rand(100) * 0.1, torch.randn(100) * 4
        p0, m0, v0 = p.clone(), m.clone(), v.clone()
        args = (1e-2, 0.9, 0.999, 1e-8, 2.0, 3, mode, 1, 0.01)
        copy = torch.empty(100, dtype=torch.bfloat16)
        F.reversible_adam(p, copy, m, v, g, *args)
        assert not torch.equal(p, p0)
        torch.testing.assert_close(copy.float(), p, atol=2e-2, rtol=2e-2)
        F.maybe_adam_undo(torch.zeros(1), p, m, v, g, *args)           # flag clear: nothing happens
        assert not torch.equal(p, p0)
        F.maybe_adam_undo(torch.ones(1), p, m, v, g, *args)
        for a, b in ((p, p0), (m, m0), (v, v0)):
            torch.testing.assert_close(a, b, atol=1e-6, rtol=1e-5)
    flag, x = torch.zeros(1), torch.ones(10)
    x[4] = float("inf")
    F.strided_check_finite(flag, x, 2, 1)
    assert flag.item() == 1
    F.strided_check_finite(flag, x, 3, 1)
    assert flag.item() == 0


def test_sparse_masks_property():
    """2:4 masks for random shapes: 1-D pattern keeps exactly the two largest magnitudes of every group of four along the input dimension;
    the 2-D patterns keep at most two per row AND per column of every 4x4 block; `best` (exhaustive over the 90 exact patterns) keeps exactly
    two and is never beaten by another exact pattern — greedy may stop at a maximal mask with fewer entries, which is not comparable."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from apex_b200.contrib.sparsity import sparse_masklib as L

    @settings(max_examples=25, deadline=None)
    @given(rows=st.integers(1, 6), cols=st.integers(1, 6), seed=st.integers(0, 10_000))
    def check(rows, cols, seed):
        w = torch.randn(4 * rows, 4 * cols, generator=torch.Generator().manual_seed(seed))
        m1 = L.create_mask(w, "m4n2_1d").bool()
        g, k = w.abs().view(-1, 4), m1.view(-1, 4)
        assert bool((k.sum(1) == 2).all())
        kept_min = torch.where(k, g, torch.full_like(g, float("inf"))).min(1).values
        dropped_max = torch.where(k, torch.full_like(g, -1.0), g).max(1).values
        assert bool((kept_min >= dropped_max).all())
        best, greedy = L.create_mask(w, "m4n2_2d_best").bool(), L.create_mask(w, "m4n2_2d_greedy").bool()
        for m2 in (best, greedy):
            blocks = m2.view(rows, 4, cols, 4).permute(0, 2, 1, 3)
            assert int(blocks.sum(-1).max()) <= 2 and int(blocks.sum(-2).max()) <= 2
        bb = best.view(rows, 4, cols, 4).permute(0, 2, 1, 3)
        assert bool((bb.sum(-1) == 2).all()) and bool((bb.sum(-2) == 2).all())
        gb = greedy.view(rows, 4, cols, 4).permute(0, 2, 1, 3)
        wb = w.abs().view(rows, 4, cols, 4).permute(0, 2, 1, 3)
        exact = (gb.sum(-1) == 2).all(-1) & (gb.sum(-2) == 2).all(-1)          # blocks where greedy happens to be an exact pattern
        assert bool(((wb * bb).sum((-1, -2)) >= (wb * gb).sum((-1, -2)) - 1e-5)[exact].all())

    check()


def test_reference_helper_names_exist_and_agree():
    """Small public helpers of the reference that tests / drivers import by name."""
    import warnings

    import apex_b200
    from apex_b200.contrib.conv_bias_relu import ConvBiasReLU, ConvBiasReLU_
    from apex_b200.contrib.group_norm.group_norm import group_norm_nhwc_bprop, group_norm_nhwc_fprop
    from apex_b200.contrib.openfold import FusedAdamSWA
    from apex_b200.contrib.transducer._transducer_ref import transducer_loss_reference
    from apex_b200.multi_tensor_apply import multi_tensor_applier

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert apex_b200.check_cudnn_version_and_warn("x", 10 ** 9) is False
    multi_tensor_applier.check_avail()
    x, w, b = torch.randn(2, 3, 5, 5), torch.randn(4, 3, 3, 3), torch.randn(1, 4, 1, 1)
    torch.testing.assert_close(ConvBiasReLU_.apply(x, w, b, 1, 1), ConvBiasReLU(x, w, b, 1, 1))
    # GroupNorm fprop / bprop pair against autograd through torch's group_norm
    xg = torch.randn(2, 8, 3, 3).contiguous(memory_format=torch.channels_last)
    gw, gb = torch.randn(8), torch.randn(8)
    y, sums = group_norm_nhwc_fprop(xg, 4, gw, gb, 1e-5, "silu")
    leaves = [t.clone().requires_grad_() for t in (xg, gw, gb)]
    ref = torch.nn.functional.silu(torch.nn.functional.group_norm(leaves[0], 4, leaves[1], leaves[2], 1e-5))
    torch.testing.assert_close(y, ref.detach())
    dy = torch.randn_like(y)
    for got, want in zip(group_norm_nhwc_bprop(dy, sums, xg, 4, gw, gb, 1e-5, "silu"), torch.autograd.grad(ref, leaves, dy)):
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    # transducer lattice oracle: loss == -beta[0, 0] == -(alpha[end] + blank[end])
    B, T, U, V = 2, 4, 3, 5
    lg = torch.randn(B, T, U, V, requires_grad=True)
    label, f_len, y_len = torch.randint(1, V, (B, U - 1)), torch.tensor([4, 3]), torch.tensor([2, 1])
    alpha, beta, grad, loss = transducer_loss_reference(lg, label, f_len, y_len, 0, torch.ones(B))
    torch.testing.assert_close(loss, -beta[:, 0, 0])
    end = torch.stack([alpha[i, f_len[i] - 1, y_len[i]] + torch.log_softmax(lg.detach()[i, f_len[i] - 1, y_len[i]], -1)[0] for i in range(B)])
    torch.testing.assert_close(loss, -end)
    assert grad.shape == lg.shape
    # FusedAdamSWA.from_optim continues a torch Adam run exactly
    fp32 = [torch.nn.Parameter(torch.randn(10))]
    ref_p = [torch.nn.Parameter(fp32[0].detach().clone())]
    adam, adam_ref = torch.optim.Adam(fp32, lr=1e-2), torch.optim.Adam(ref_p, lr=1e-2)
    for it in range(4):
        g = torch.randn(10, generator=torch.Generator().manual_seed(it))
        if it == 2:
            bf = [torch.nn.Parameter(fp32[0].detach().bfloat16())]
            opt = FusedAdamSWA.from_optim(adam, fp32, bf, [torch.nn.Parameter(fp32[0].detach().clone())], 0.9)
        if it < 2:
            fp32[0].grad = g.clone()
            adam.step()
        else:
            bf[0].grad = g.bfloat16()
            g = bf[0].grad.float()
            opt.step()
        ref_p[0].grad = g.clone()
        adam_ref.step()
    torch.testing.assert_close(fp32[0], ref_p[0])


def test_every_submodule_imports():
    import importlib
    import pkgutil

    import apex_b200

    failures = []
    for m in pkgutil.walk_packages(apex_b200.__path__, "apex_b200."):
        if any(part in m.name for part in ("._C", "._build", "._kernels", ".csrc", "__main__")):
            continue
        try:
            importlib.import_module(m.name)
        except Exception as e:  # noqa: BLE001
            failures.append((m.name, repr(e)))
    assert not failures, failures


def test_contrib_raw_extension_names_cpu(tmp_path):
    """The contrib extension names registered by ext_compat (fused_conv_bias_relu, group_norm_cuda / _v2_cuda, _apex_gpu_direct_storage,
    nccl_p2p_cuda, _apex_nccl_allocator, transducer_*_cuda) with the reference's raw calling conventions; CPU tensors take the library's
    PyTorch paths, so the list / tuple conventions and argument orders are what is checked here."""
    from apex_b200 import ext_compat as E
    m = E.extension_modules()
    for name in ("fused_conv_bias_relu", "group_norm_cuda", "group_norm_v2_cuda", "transducer_joint_cuda", "transducer_loss_cuda", "nccl_p2p_cuda",
                 "_apex_nccl_allocator", "_apex_gpu_direct_storage"):
        assert name in m, name
    torch.manual_seed(0)
    cb = m["fused_conv_bias_relu"]
    x = torch.randn(2, 8, 6, 6).contiguous(memory_format=torch.channels_last)
    w = torch.randn(16, 8, 3, 3).contiguous(memory_format=torch.channels_last)
    b = torch.randn(1, 16, 1, 1)
    out = cb.forward([x, w, b], 1, 1)[0]
    torch.testing.assert_close(out, torch.relu(torch.nn.functional.conv2d(x, w, b.reshape(-1), 1, 1)))
    dy = torch.randn_like(out)
    dx, dw, db = cb.backward([x, w, out, dy], 1, 1)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    torch.relu(torch.nn.functional.conv2d(xr, wr, br.reshape(-1), 1, 1)).backward(dy)
    torch.testing.assert_close(dx, xr.grad)
    torch.testing.assert_close(dw, wr.grad, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(db, br.grad, atol=1e-4, rtol=1e-4)
    out2 = cb.forward_no_relu([x, w, b], 1, 1)[0]
    torch.testing.assert_close(out2, torch.nn.functional.conv2d(x, w, b.reshape(-1), 1, 1))
    sc = torch.rand(1, 16, 1, 1) + 0.5
    out3 = cb.forward_cscale_cbias_relu([x, w, sc, b], 1, 1)[0]
    torch.testing.assert_close(out3, torch.relu(torch.nn.functional.conv2d(x, w, None, 1, 1) * sc + b))
    g = cb.backward_cscale_cbias_relu([x, w, sc, out3, dy], 1, 1)
    assert len(g) == 2 and g[0].shape == x.shape and g[1].shape == w.shape

    gn = m["group_norm_cuda"]
    xg = torch.randn(2, 8, 4, 4).contiguous(memory_format=torch.channels_last)
    wg, bg = torch.randn(8), torch.randn(8)
    y, sums = gn.forward(xg, 4, wg, bg, 1e-5, 1, True)
    torch.testing.assert_close(y, torch.nn.functional.silu(torch.nn.functional.group_norm(xg, 4, wg, bg, 1e-5)))
    assert sums.shape == (2 * 2 * 4,)
    xr = xg.clone().requires_grad_(True)
    wr, br2 = wg.clone().requires_grad_(True), bg.clone().requires_grad_(True)
    torch.nn.functional.silu(torch.nn.functional.group_norm(xr, 4, wr, br2, 1e-5)).backward(torch.ones_like(y))
    dxg, dwg, dbg = gn.backward(torch.ones_like(y), sums, xg, 4, wg, bg, 1e-5, 1, True)
    torch.testing.assert_close(dxg, xr.grad, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(dwg, wr.grad, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(dbg, br2.grad, atol=1e-4, rtol=1e-4)
    mv = torch.empty(16)
    y2 = m["group_norm_v2_cuda"].gn(xg, wg, bg, 1e-5, True, 4, mean_var_out=mv)
    torch.testing.assert_close(y2, y)
    torch.testing.assert_close(mv, sums)
    d2 = m["group_norm_v2_cuda"].gn_bwd(torch.ones_like(y), xg, wg, bg, mv, 1e-5, True, 4)
    torch.testing.assert_close(d2[0], dxg)

    F_ = m["_apex_gpu_direct_storage"]._GDSFile
    path = str(tmp_path / "t.bin")
    f = F_(path, "w"); t = torch.arange(10.); f.save_data(t); f.close()
    f = F_(path, "r"); u = torch.empty(10); f.load_data(u); f.close()
    assert torch.equal(t, u)
    assert callable(m["nccl_p2p_cuda"].get_unique_nccl_id) and callable(m["_apex_nccl_allocator"].get_nccl_allocator)


def test_permutation_search_cuda_raw_entry_points():
    """``permutation_search_cuda`` with the reference's numpy-buffer conventions (results written into the caller's arrays)."""
    import numpy as np
    from apex_b200 import ext_compat as E
    from apex_b200.contrib.sparsity.permutation_search import generate_all_unique_combinations
    m = E.extension_modules()["permutation_search_cuda"]
    rng = np.random.default_rng(0)
    M = rng.standard_normal((16, 16)).astype(np.float32)

    def kept(x):
        return float(sum(np.sort(np.abs(x[r, c:c + 4]))[2:].sum() for r in range(x.shape[0]) for c in range(0, x.shape[1], 4)))

    out = np.zeros(1, dtype=np.float32)
    assert m.sum_after_2_to_4(M.flatten(), 16, 16, 0, 16, 2, 4, out) == 0
    assert abs(out[0] - kept(M)) < 1e-3
    m.sum_after_2_to_4(M.flatten(), 16, 16, 4, 12, 2, 4, out)
    assert abs(out[0] - kept(M[:, 4:12])) < 1e-3
    perms = generate_all_unique_combinations(8, 4)
    imp, idx = np.zeros(2, dtype=np.float32), np.zeros(2, dtype=np.uint32)
    m.build_permute_map(M.flatten(), 16, 16, np.array([0, 1, 2, 3], dtype=np.uint32), 2, 2, perms.astype(np.uint32).flatten(), 8, imp, idx)
    for g, lo in enumerate((0, 8)):
        sub = M[:, lo:lo + 8]
        assert abs(imp[g] - (max(kept(sub[:, p]) for p in perms) - kept(sub))) < 1e-3
        assert abs(kept(sub[:, perms[idx[g]]]) - kept(sub) - imp[g]) < 1e-3
    imp1, pi = np.zeros(1, dtype=np.float32), np.zeros(1, dtype=np.uint32)
    m.check_permutations(M.flatten(), 16, 16, np.array([0, 1], dtype=np.uint32), 2, 1, perms.astype(np.uint32).flatten(), len(perms), imp1, pi)
    assert abs(imp1[0] - imp[0]) < 1e-3 and pi[0] == idx[0]
    o = np.zeros(16, dtype=np.float32)
    m.build_swap_map(M.flatten(), 16, 16, np.array([0, 1], dtype=np.uint32), o)
    for k in (0, 5, 15):
        c = M[:, :8].copy()
        a, b = k // 4, 4 + k % 4
        c[:, [a, b]] = c[:, [b, a]]
        assert abs(o[k] - (kept(c) - kept(M[:, :8]))) < 1e-3


def test_cudnn_gbn_lib_single_rank_matches_batch_norm():
    from apex_b200 import ext_compat as E
    g = E.extension_modules()["cudnn_gbn_lib"]
    torch.manual_seed(0)
    x = torch.randn(4, 8, 5, 5).contiguous(memory_format=torch.channels_last)
    w, b = torch.randn(8), torch.randn(8)
    rm, rv, mm, miv = torch.zeros(8), torch.ones(8), torch.empty(8), torch.empty(8)
    y = g.forward(x, w, b, rm, rv, mm, miv, 0.1, 1e-5, 1, 0, [])
    rm2, rv2 = torch.zeros(8), torch.ones(8)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.batch_norm(xr, rm2, rv2, wr, br, True, 0.1, 1e-5)
    torch.testing.assert_close(y, ref.detach(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(rm, rm2)
    torch.testing.assert_close(rv, rv2, atol=1e-6, rtol=1e-5)
    dy = torch.randn_like(y)
    ref.backward(dy)
    dx, dw, db = g.backward(x, dy, w, mm, miv, 1e-5, 1, 0, [])
    torch.testing.assert_close(dx, xr.grad, atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(dw, wr.grad, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(db, br.grad, atol=1e-4, rtol=1e-4)


def test_peer_memory_cuda_extension_name_cpu():
    """Registration, the typed-view stride rule (reference peer_memory_cuda.cu:34-55) and the loud failure without a GPU."""
    import pytest
    import torch
    from apex_b200 import ext_compat
    m = ext_compat.extension_modules()["peer_memory_cuda"]
    for n in ("allocate_raw", "free_raw", "zero", "get_raw_ipc_address", "get_raw_peers", "blob_view_half", "blob_view_float",
              "blob_view_int", "push_pull_halos_1d"):
        assert callable(getattr(m, n))
    assert ext_compat.blob_strides([2, 3, 4, 5], False) == list(torch.empty(2, 3, 4, 5).stride())
    assert ext_compat.blob_strides([2, 3, 4, 5], True) == list(torch.empty(2, 3, 4, 5).contiguous(memory_format=torch.channels_last).stride())
    assert ext_compat.blob_strides([7], False) == [1]
    with pytest.raises(ValueError):
        m.get_raw_ipc_address(1234)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA device"):
            m.allocate_raw(1024)


def test_fast_bottleneck_extension_name_cpu():
    """fast_bottleneck.forward / backward on explicit tensor lists, driven the way the reference's BottleneckFunction drives them
    (bottleneck.py:80-132), against autograd through the same block written with plain torch ops."""
    import pytest
    import torch
    import torch.nn.functional as F
    from apex_b200 import ext_compat
    fb = ext_compat.extension_modules()["fast_bottleneck"]
    for nhwc, stride, down in ((False, 1, False), (True, 2, True), (False, 2, True)):
        torch.manual_seed(0)
        Cin, Cmid, Cout = (8, 4, 8) if not down else (8, 4, 16)
        x = torch.randn(2, Cin, 8, 8, dtype=torch.double, requires_grad=True)
        ws = [torch.randn(Cmid, Cin, 1, 1, dtype=torch.double, requires_grad=True), torch.randn(Cmid, Cmid, 3, 3, dtype=torch.double, requires_grad=True),
              torch.randn(Cout, Cmid, 1, 1, dtype=torch.double, requires_grad=True)]
        if down:
            ws.append(torch.randn(Cout, Cin, 1, 1, dtype=torch.double, requires_grad=True))
        ss = [torch.rand(w.shape[0], dtype=torch.double) + 0.5 for w in ws]
        bs = [torch.randn(w.shape[0], dtype=torch.double) for w in ws]
        v = lambda t: t.view(1, -1, 1, 1)
        o1 = torch.relu(F.conv2d(x, ws[0], stride=stride) * v(ss[0]) + v(bs[0]))
        o2 = torch.relu(F.conv2d(o1, ws[1], padding=1) * v(ss[1]) + v(bs[1]))
        idn = F.conv2d(x, ws[3], stride=stride) * v(ss[3]) + v(bs[3]) if down else x
        o3 = torch.relu(F.conv2d(o2, ws[2]) * v(ss[2]) + v(bs[2]) + idn)
        lay = (lambda t: t.detach().permute(0, 2, 3, 1).contiguous()) if nhwc else (lambda t: t.detach())
        args = [lay(x)] + [lay(w) for w in ws[:3]] + ss[:3] + bs[:3] + ([lay(ws[3]), ss[3], bs[3]] if down else [])
        outs = fb.forward(nhwc, stride, args)
        for got, want in zip(outs, (o1, o2, o3)):
            torch.testing.assert_close(got, lay(want))
        go = torch.randn_like(o3)
        want = torch.autograd.grad(o3, [x] + ws, go)
        dre = go * (o3 > 0)
        t_list = args[:10] + [lay(dre * v(ss[2])), lay(dre * v(ss[3]) if down else dre), outs[0], outs[1]] + ([args[10]] if down else [])
        got = fb.backward(nhwc, stride, t_list)
        assert len(got) == len(want)
        for a, r in zip(got, want):
            torch.testing.assert_close(a, lay(r))


def test_bnp_extension_name_cpu():
    """Raw bnp entry points (contrib/groupbn/raw_ext.py), bn_group = 1: forward state tensors, running statistics, the ReLU bitmask and both
    backward calls against autograd; the grouped case runs in tests/test_cpu_ddp.py::test_group_batchnorm_four_ranks_gloo."""
    import torch
    import torch.nn.functional as F
    from apex_b200 import ext_compat
    from apex_b200.contrib.groupbn import raw_ext
    bnp = ext_compat.extension_modules()["bnp"]
    assert all(callable(getattr(bnp, n)) for n in raw_ext.ENTRY_POINTS)
    torch.manual_seed(0)
    N, H, W, C = 3, 5, 4, 6
    x, z = torch.randn(N, H, W, C, requires_grad=True), torch.randn(N, H, W, C, requires_grad=True)
    w, b = (torch.rand(C) + 0.5).requires_grad_(True), torch.randn(C, requires_grad=True)
    tail = (None, None, None, None, 1, torch.IntTensor([0]), 2, 100, False)
    bits = torch.randint(0, 2, (1000,)).bool()
    packed = torch.zeros(40, dtype=torch.int32)
    raw_ext._pack_bits(bits, packed)
    assert torch.equal(raw_ext._unpack_bits(packed, (1000,)), bits)
    for relu in (False, True):
        rm, rv, rm2, rv2 = torch.zeros(C), torch.ones(C), torch.zeros(C), torch.ones(C)
        mm, mi = torch.empty(C), torch.empty(C)
        y = bnp.bn_fwd_nhwc(x.detach(), w.detach(), b.detach(), rm, rv, mm, mi, None, 0.1, 1e-5, relu, *tail)
        ref = F.batch_norm(x.permute(0, 3, 1, 2), rm2, rv2, w, b, True, 0.1, 1e-5).permute(0, 2, 3, 1)
        ref = torch.relu(ref) if relu else ref
        torch.testing.assert_close(y, ref, atol=1e-5, rtol=1e-5)
        torch.testing.assert_close(rm, rm2, atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(rv, rv2, atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(mm, x.detach().mean((0, 1, 2)), atol=1e-6, rtol=1e-5)
        gy = torch.randn_like(y)
        got = bnp.bn_bwd_nhwc(x.detach(), gy, w.detach(), b.detach(), rm, rv, mm, mi, None, 0.1, 1e-5, relu, *tail)
        for a, r in zip(got, torch.autograd.grad(ref, (x, w, b), gy)):
            torch.testing.assert_close(a, r, atol=1e-4, rtol=1e-4)
        ye = bnp.bn_fwd_eval_nhwc(x.detach(), w.detach(), b.detach(), rm, rv, None, 1, 0.1, 1e-5, relu)
        re = F.batch_norm(x.detach().permute(0, 3, 1, 2), rm, rv, w.detach(), b.detach(), False, 0.0, 1e-5).permute(0, 2, 3, 1)
        torch.testing.assert_close(ye, torch.relu(re) if relu else re, atol=1e-5, rtol=1e-5)
    rm, rv, mm, mi = torch.zeros(C), torch.ones(C), torch.empty(C), torch.empty(C)
    bitmask = torch.zeros(((x.numel() + 31) // 32) * 2, dtype=torch.int32)
    y = bnp.bn_addrelu_fwd_nhwc(x.detach(), z.detach(), w.detach(), b.detach(), rm, rv, mm, mi, bitmask, None, 0.1, 1e-5, *tail)
    ref = torch.relu(F.batch_norm(x.permute(0, 3, 1, 2), None, None, w, b, True, 0.1, 1e-5).permute(0, 2, 3, 1) + z)
    torch.testing.assert_close(y, ref, atol=1e-5, rtol=1e-5)
    gy = torch.randn_like(y)
    got = bnp.bn_addrelu_bwd_nhwc(x.detach(), gy, w.detach(), b.detach(), rm, rv, mm, mi, bitmask, None, 0.1, 1e-5, *tail)
    for a, r in zip(got, torch.autograd.grad(ref, (x, z, w, b), gy)):
        torch.testing.assert_close(a, r, atol=1e-4, rtol=1e-4)
    ye = bnp.bn_addrelu_fwd_eval_nhwc(x.detach(), z.detach(), w.detach(), b.detach(), rm, rv, None, 1, 0.1, 1e-5)
    re = F.batch_norm(x.detach().permute(0, 3, 1, 2), rm, rv, w.detach(), b.detach(), False, 0.0, 1e-5).permute(0, 2, 3, 1) + z.detach()
    torch.testing.assert_close(ye, torch.relu(re), atol=1e-5, rtol=1e-5)


def test_nccl_allocator_symmetric_keyword_selection_cpu():
    import pytest
    from apex_b200.contrib.nccl_allocator import nccl_allocator as na

    class Upstream:
        def __init__(self, allocator=None, symmetric=False): ...

    class Nvidia:
        def __init__(self, allocator=None, symm_mem=False): ...

    class Old:
        def __init__(self, allocator=None): ...

    assert na.get_func_args(lambda a, b=1, *c, **d: 0) == ["a", "b", "c", "d"]
    assert na._symmetric_kwargs(None, Old) == {}
    assert na._symmetric_kwargs(True, Upstream) == {"symmetric": True}
    assert na._symmetric_kwargs(False, Nvidia) == {"symm_mem": False}
    with pytest.raises(ValueError, match="higher PyTorch version"):
        na._symmetric_kwargs(True, Old)
