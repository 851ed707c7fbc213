"""Test bodies that run inside spawned ranks (must be importable by name from child processes)."""
import copy

import torch
import torch.distributed as dist


def _model(device, dtype=torch.float32, seed=0):
    torch.manual_seed(seed)
    # 11 small layers whose sizes straddle bucket boundaries (reference test_dist_adam.py:47 uses a 71-element bucket)
    m = torch.nn.Sequential(*[torch.nn.Linear(7, 7) for _ in range(11)]).to(device=device, dtype=dtype)
    return m


def dist_adam_matches_ddp_adamw(rank, world, device_type, fused, steps=4, clip=False, dtype=torch.float32, grad_sync_dtype=None, param_sync_dtype=None):
    """Oracle = every rank holds the full model + torch.optim.AdamW on all-reduced (averaged) grads."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    ref_model = _model(dev)
    dist_model = copy.deepcopy(ref_model).to(dtype)
    groups = lambda m: [{"params": [p for n, p in m.named_parameters() if "weight" in n], "lr": 3e-3},
                        {"params": [p for n, p in m.named_parameters() if "bias" in n], "lr": 1e-2, "weight_decay": 0.0}]
    ref_opt = torch.optim.AdamW(groups(ref_model), lr=3e-3, weight_decay=0.05)
    opt = DistributedFusedAdam(groups(dist_model), lr=3e-3, weight_decay=0.05, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20,
                               fused_collectives=("auto" if fused else False), grad_sync_dtype=grad_sync_dtype, param_sync_dtype=param_sync_dtype)
    assert opt.fused_collectives == (fused and device_type == "cuda")
    g = torch.Generator().manual_seed(100 + rank)
    for it in range(steps):
        opt.zero_grad()
        ref_opt.zero_grad()
        for micro in range(2):
            x = torch.randn(5, 7, generator=g).to(dev)
            y_ref = ref_model(x).pow(2).mean()
            y_ref.backward()
            with opt.no_sync():
                dist_model(x.to(dtype)).float().pow(2).mean().backward()
        for p in ref_model.parameters():
            dist.all_reduce(p.grad)
            p.grad /= world
        if clip:
            n_ref = torch.nn.utils.clip_grad_norm_(ref_model.parameters(), 0.05)
            n = opt.clip_grad_norm(0.05)
            torch.testing.assert_close(n.cpu(), n_ref.cpu(), rtol=2e-2 if dtype != torch.float32 else 1e-4, atol=1e-5)
        ref_opt.step()
        opt.step()
        tol = 1e-5 if (dtype == torch.float32 and grad_sync_dtype is None) else 3e-2
        for pr, pd in zip(ref_model.parameters(), dist_model.parameters()):
            torch.testing.assert_close(pd.float(), pr, rtol=tol, atol=tol)
    # every rank ends with identical parameters
    for p in dist_model.parameters():
        t = p.detach().float().clone()
        dist.broadcast(t, src=0)
        torch.testing.assert_close(t, p.detach().float(), rtol=0, atol=0)


def dist_adam_overlap_grad_sync(rank, world, device_type):
    """overlap_grad_sync: the last micro-batch runs outside no_sync(), so post-accumulate-grad hooks reduce-scatter every bucket on the side
    stream while backward is still running; the result must equal DDP + AdamW with micro-batch accumulation
    (reference apex/contrib/test/optimizers/test_dist_adam.py:119-300)."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    # six 48x48 layers: ~14k elements over buckets of 2048 * world elements => several buckets, some parameters straddle two
    ref_model = torch.nn.Sequential(*[torch.nn.Linear(48, 48) for _ in range(6)]).to(dev)
    dist_model = copy.deepcopy(ref_model)
    ref_opt = torch.optim.AdamW(ref_model.parameters(), lr=3e-3, weight_decay=0.05)
    opt = DistributedFusedAdam(dist_model.parameters(), lr=3e-3, weight_decay=0.05, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20,
                               overlap_grad_sync=True)
    assert opt.fused_collectives
    g = torch.Generator().manual_seed(100 + rank)
    for it in range(4):
        opt.zero_grad(set_to_none=bool(it & 1))      # both gradient paths: in-place accumulation and single-copy hand-over
        ref_opt.zero_grad()
        for micro in range(3):
            x = torch.randn(5, 48, generator=g).to(dev)
            ref_model(x).pow(2).mean().backward()
            if micro < 2:
                with opt.no_sync():
                    dist_model(x).pow(2).mean().backward()
            else:
                dist_model(x).pow(2).mean().backward()
        assert any(any(seg.bucket_synced) for seg in opt._segments), "no bucket was reduce-scattered during backward"
        for p in ref_model.parameters():
            dist.all_reduce(p.grad)
            p.grad /= world
        if it % 2:
            n_ref = torch.nn.utils.clip_grad_norm_(ref_model.parameters(), 0.05)
            torch.testing.assert_close(opt.clip_grad_norm(0.05).cpu(), n_ref.cpu(), rtol=1e-4, atol=1e-5)
        ref_opt.step()
        opt.step()
        for pr, pd in zip(ref_model.parameters(), dist_model.parameters()):
            torch.testing.assert_close(pd, pr, rtol=1e-5, atol=1e-5)
    n_buckets = sum(seg.n_buckets for seg in opt._segments)
    assert n_buckets > 1 and opt.kernel_launches >= 4 * (n_buckets + 1)


def dist_adam_step_in_backward(rank, world, device_type):
    """overlap_step_with_backward: every bucket's whole step (reduce-scatter + Adam + parameter push) runs from the gradient hook while
    backward continues; step() only joins. Same kernel on the same data => bit-identical to a twin that steps after backward."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    ma = torch.nn.Sequential(*[torch.nn.Linear(48, 48) for _ in range(6)]).to(dev)
    mb = copy.deepcopy(ma)
    kw = dict(lr=3e-3, weight_decay=0.05, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20)
    a = DistributedFusedAdam(ma.parameters(), overlap_step_with_backward=True, **kw)
    b = DistributedFusedAdam(mb.parameters(), overlap_grad_sync=False, **kw)
    g = torch.Generator().manual_seed(100 + rank)
    for it in range(5):
        a.zero_grad(set_to_none=bool(it & 1))
        b.zero_grad()
        for micro in range(2):
            x = torch.randn(5, 48, generator=g).to(dev)
            if micro == 0:
                with a.no_sync():
                    ma(x).pow(2).mean().backward()
            else:
                ma(x).pow(2).mean().backward()
            mb(x).pow(2).mean().backward()
        assert any(any(seg.bucket_stepped) for seg in a._segments), "no bucket stepped during backward"
        a.step()
        b.step()
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.equal(pa, pb), (it, (pa - pb).abs().max())
    torch.testing.assert_close(a.last_grad_norm(), b.last_grad_norm(), rtol=1e-5, atol=1e-7)
    assert a.param_groups[0]["step"] == b.param_groups[0]["step"] == 5


def dist_adam_overlap_param_sync(rank, world, device_type):
    """overlap_param_sync: step() returns while Adam + the parameter push run on the side stream; the FusedDense forward that follows
    acquires per-bucket ready flags inside its GEMM kernels (weights) and through stream-ordered waits (biases, via the module hooks).
    Must equal a twin that joins everything inside step()."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    from apex_b200.fused_dense import FusedDense
    from apex_b200.ops import gemm as G
    from apex_b200.parallel.param_sync import attach_param_sync_hooks
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    mk = lambda: torch.nn.Sequential(FusedDense(256, 512), torch.nn.GELU(), FusedDense(512, 256), torch.nn.LayerNorm(256)).to(dev, torch.bfloat16)  # noqa: E731
    ma = mk()
    mb = copy.deepcopy(ma)
    kw = dict(lr=1e-2, weight_decay=0.01, device=dev, bucket_cap_mb=0.1)
    a = DistributedFusedAdam(ma.parameters(), overlap_param_sync=True, **kw)
    # the twin gets its own process group => its own signal pad / epoch counters: its step kernels run on the main stream WHILE a's
    # parameter push is still in flight on the side stream (one optimizer never overlaps two of its own collectives)
    b = DistributedFusedAdam(mb.parameters(), overlap_grad_sync=False, process_group=dist.new_group(list(range(world))), **kw)
    attach_param_sync_hooks(ma)
    g = torch.Generator().manual_seed(100 + rank)
    guarded0 = G.stats["guarded"]
    for it in range(5):
        x = torch.randn(64, 256, generator=g).to(dev, torch.bfloat16)
        a.zero_grad()
        b.zero_grad()
        ya = ma(x)                     # reads weights that the previous step may still be pushing
        yb = mb(x)
        assert torch.equal(ya, yb), (it, (ya.float() - yb.float()).abs().max())
        ya.float().pow(2).mean().backward()
        yb.float().pow(2).mean().backward()
        a.step()
        b.step()
    a.param_sync()
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert torch.equal(pa, pb)
    assert G.stats["guarded"] - guarded0 >= 4, "the forward GEMMs never saw an in-flight parameter buffer"


def dist_adam_cuda_graph_replays(rank, world, device_type):
    """capturable=True at D > 1: torch.cuda.graph capture of step() + 10 replays must match an eager twin bit for bit — the collective's
    epoch lives in device memory, so every replay signals / waits on a fresh value (reference test_dist_adam.py:834)."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank)
    torch.manual_seed(7)
    shapes = [(257, 33), (4099,), (64, 64)]
    mk = lambda: [torch.nn.Parameter(torch.randn(s, device=dev).bfloat16()) for s in shapes]  # noqa: E731
    torch.manual_seed(7)
    pa = mk()
    torch.manual_seed(7)
    pb = mk()
    kw = dict(lr=1e-2, weight_decay=0.01, device=dev, capturable=True, bucket_cap_mb=0.05, overlap_grad_sync=False)
    a, b = DistributedFusedAdam(pa, **kw), DistributedFusedAdam(pb, **kw)
    a.zero_grad()
    b.zero_grad()
    gen = torch.Generator(device=dev).manual_seed(50 + rank)

    def new_grads():
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, device=dev, generator=gen).bfloat16()
            x.grad.copy_(gr)
            y.grad.copy_(gr)

    for _ in range(2):   # warm-up (allocations, lazy init) outside the capture
        new_grads()
        a.step()
        b.step()
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    new_grads()
    with torch.cuda.graph(graph):
        a.step()
    b.step()        # the capture does not execute: a is one step behind until the first replay
    graph.replay()
    for _ in range(10):
        new_grads()
        graph.replay()
        b.step()
    torch.cuda.synchronize()
    for x, y in zip(pa, pb):
        assert torch.equal(x, y)
    for sa, sb in zip(a._segments, b._segments):
        assert torch.equal(sa.master, sb.master) and torch.equal(sa.exp_avg_sq, sb.exp_avg_sq)
    assert int(a.param_groups[0]["step"].item()) == int(b.param_groups[0]["step"].item()) == 13


def dist_adam_two_dimensional_grid(rank, world, device_type):
    """distributed x redundant process-group grid (the reference's HSDP-like layout, distributed_fused_adam.py:477-560): state is sharded
    inside each distributed group of 2 and replicated across the redundant groups; 4 ranks must still equal data-parallel AdamW over all 4."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    assert world == 4
    dev = torch.device("cpu")
    dgroups = [dist.new_group([0, 1]), dist.new_group([2, 3])]
    rgroups = [dist.new_group([0, 2]), dist.new_group([1, 3])]
    dg, rg = dgroups[rank // 2], rgroups[rank % 2]
    ref_model = _model(dev)
    model = copy.deepcopy(ref_model)
    ref_opt = torch.optim.AdamW(ref_model.parameters(), lr=3e-3, weight_decay=0.05)
    opt = DistributedFusedAdam(model.parameters(), lr=3e-3, weight_decay=0.05, device=dev, distributed_process_group=dg, redundant_process_group=rg,
                               bucket_cap_mb=2048 * 4 * 2 / 2 ** 20)
    assert (opt.distributed_size, opt.redundant_size) == (2, 2)
    g = torch.Generator().manual_seed(100 + rank)
    for it in range(3):
        opt.zero_grad()
        ref_opt.zero_grad()
        x = torch.randn(5, 7, generator=g)
        ref_model(x).pow(2).mean().backward()
        model(x).pow(2).mean().backward()
        for p in ref_model.parameters():
            dist.all_reduce(p.grad)
            p.grad /= world
        if it == 1:
            torch.testing.assert_close(opt.clip_grad_norm(0.05), torch.nn.utils.clip_grad_norm_(ref_model.parameters(), 0.05), rtol=1e-4, atol=1e-6)
        ref_opt.step()
        opt.step()
        for pr, pd in zip(ref_model.parameters(), model.parameters()):
            torch.testing.assert_close(pd, pr, rtol=1e-5, atol=1e-5)


def dist_adam_scaled_states_on_several_ranks(rank, world, device_type):
    """with_scaled_states (bf16 state + per-fragment fp32 scales) sharded over ranks, buckets that parameters straddle: stays finite, tracks
    data-parallel fp32 AdamW to bf16 accuracy, every rank ends with identical parameters (reference test_matches_pytorch_scaled_state)."""
    import warnings

    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    ref = _model(dev)
    model = copy.deepcopy(ref).bfloat16()
    ref_opt = torch.optim.AdamW(ref.parameters(), lr=3e-3, weight_decay=0.05)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt = DistributedFusedAdam(model.parameters(), lr=3e-3, weight_decay=0.05, device=dev, dtype=torch.bfloat16, with_scaled_states=True,
                                   bucket_cap_mb=2048 * 4 * world / 2 ** 20, fused_collectives=False)
    g = torch.Generator().manual_seed(100 + rank)
    for it in range(4):
        opt.zero_grad()
        ref_opt.zero_grad()
        x = torch.randn(5, 7, generator=g).to(dev)
        ref(x).pow(2).mean().backward()
        model(x.bfloat16()).float().pow(2).mean().backward()
        for p in ref.parameters():
            dist.all_reduce(p.grad)
            p.grad /= world
        ref_opt.step()
        opt.step()
    for a, b in zip(model.parameters(), ref.parameters()):
        assert bool(torch.isfinite(a.float()).all())
        torch.testing.assert_close(a.float(), b, rtol=5e-2, atol=5e-2)
        t = a.detach().float().clone()
        dist.broadcast(t, src=0)
        torch.testing.assert_close(t, a.detach().float(), rtol=0, atol=0)


def dist_adam_state_dict_reshards(rank, world, device_type, tmpdir):
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    model = _model(dev)
    opt = DistributedFusedAdam(model.parameters(), lr=1e-2, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20, fused_collectives=False)
    g = torch.Generator().manual_seed(7)
    def one_step(m, o):
        o.zero_grad()
        x = torch.randn(5, 7, generator=g).to(dev)
        m(x).pow(2).mean().backward()
        o.step()
    one_step(model, opt)
    sd = opt.state_dict()
    if rank == 0:
        torch.save({"opt": sd, "model": model.state_dict()}, f"{tmpdir}/ck.pt")
    dist.barrier()
    # reload on a "different world": a single-rank (no process group) optimizer on rank 0, and compare the next step
    ck = torch.load(f"{tmpdir}/ck.pt", weights_only=False)
    if rank == 0:
        m1 = _model(dev)
        m1.load_state_dict(ck["model"])
        solo_pg = dist.new_group(ranks=[0])
    else:
        solo_pg = dist.new_group(ranks=[0])
    g_state = g.get_state()
    one_step(model, opt)
    if rank == 0:
        o1 = DistributedFusedAdam(m1.parameters(), lr=1e-2, device=dev, process_group=solo_pg, fused_collectives=False)
        o1.load_state_dict(ck["opt"])
        # same data, but the solo optimizer sees only rank 0's gradient: feed it the averaged gradient explicitly
        g.set_state(g_state)
    # (numerical continuation is checked in the single-process test; here the contract is: loads without error, same keys)
    assert set(sd["state"].keys()) == set(range(22)) | {"step"}
    assert sd["state"][0]["exp_avg"].shape == (7, 7)


def dist_adam_checkpoint_moves_between_world_sizes(rank, world, device_type, layout):
    """Save on ``world`` ranks, load into a ONE-rank optimizer and into a fresh ``world``-rank optimizer, continue all three on the same data
    (every rank sees identical batches, so the averaged gradient is the single-rank gradient): parameters must stay identical
    (reference test_dist_adam.py::test_checkpoint_save_1gpu / test_checkpoint_load_1gpu, for the fp32-master and parameter-remainder layouts)."""
    import warnings

    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    dt = torch.bfloat16 if layout == "remainders" else torch.float32
    kw = dict(dtype=torch.float32, grad_sync_dtype=torch.float32, param_sync_dtype=torch.bfloat16, store_params=False, store_param_remainders=True) \
        if layout == "remainders" else {}
    solo_pg = dist.new_group(ranks=[0])

    def make(pg, cap_world):
        m = _model(dev).to(dt)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o = DistributedFusedAdam(m.parameters(), lr=1e-2, weight_decay=0.05, device=dev, process_group=pg,
                                     bucket_cap_mb=2048 * 4 * cap_world / 2 ** 20, fused_collectives=False, **kw)
        return m, o

    def steps(m, o, seeds):
        for sd_ in seeds:
            o.zero_grad()
            x = torch.randn(5, 7, generator=torch.Generator().manual_seed(sd_)).to(dev, dt)
            m(x).float().pow(2).mean().backward()
            o.step()

    model, opt = make(None, world)
    steps(model, opt, range(3))
    ck = opt.state_dict()                         # identical on every rank, independent of the world size
    again, opt_again = make(None, world)
    opt_again.load_state_dict(ck)
    steps(model, opt, range(3, 6))
    steps(again, opt_again, range(3, 6))
    for a, b in zip(again.parameters(), model.parameters()):
        torch.testing.assert_close(a, b, rtol=0, atol=0)
    if rank == 0:
        solo, opt_solo = make(solo_pg, 1)
        opt_solo.load_state_dict(ck)
        steps(solo, opt_solo, range(3, 6))
        for a, b in zip(solo.parameters(), model.parameters()):
            torch.testing.assert_close(a.float(), b.float(), rtol=1e-6, atol=1e-6)


def dist_adam_grad_scaler_skips_on_inf(rank, world, device_type):
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank)
    model = _model(dev)
    # overlap_grad_sync=False: the test plants the inf AFTER backward; with the overlap the bucket may already have been reduce-scattered
    # by then (as in the reference, gradients must not be edited behind the hooks' back)
    opt = DistributedFusedAdam(model.parameters(), lr=1e-2, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20, overlap_grad_sync=False)
    scaler = torch.amp.GradScaler("cuda", init_scale=128.0)
    x = torch.randn(5, 7, device=dev)
    before = [p.detach().clone() for p in model.parameters()]
    opt.zero_grad()
    scaler.scale(model(x).pow(2).mean()).backward()
    if rank == world - 1:
        next(model.parameters()).grad.view(-1)[3] = float("inf")
    scaler.step(opt)
    scaler.update()
    for b, p in zip(before, model.parameters()):
        torch.testing.assert_close(b, p.detach(), rtol=0, atol=0)  # skipped everywhere, including ranks that saw finite grads
    assert scaler.get_scale() == 64.0
    opt.zero_grad()
    scaler.scale(model(x).pow(2).mean()).backward()
    scaler.step(opt)
    scaler.update()
    assert any((b - p.detach()).abs().max() > 0 for b, p in zip(before, model.parameters()))


def dist_lamb_matches_fused_lamb(rank, world, device_type):
    """Oracle: FusedLAMB (same library, single process math) on the all-reduced averaged gradients."""
    from apex_b200.contrib.optimizers import DistributedFusedLAMB
    from apex_b200.optimizers import FusedLAMB
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    ref_model = _model(dev)
    dist_model = copy.deepcopy(ref_model)
    ref_opt = FusedLAMB(ref_model.parameters(), lr=2e-2, weight_decay=0.01, max_grad_norm=0.5, eps=1e-6)
    opt = DistributedFusedLAMB(dist_model.parameters(), lr=2e-2, weight_decay=0.01, max_grad_norm=0.5, eps=1e-6, device=dev,
                               bucket_cap_mb=2048 * 4 * world / 2 ** 20)
    g = torch.Generator().manual_seed(100 + rank)
    for it in range(3):
        opt.zero_grad()
        ref_opt.zero_grad()
        x = torch.randn(5, 7, generator=g).to(dev)
        ref_model(x).pow(2).mean().backward()
        dist_model(x).pow(2).mean().backward()
        for p in ref_model.parameters():
            dist.all_reduce(p.grad)
            p.grad /= world
        ref_opt.step()
        opt.step()
        for pr, pd in zip(ref_model.parameters(), dist_model.parameters()):
            torch.testing.assert_close(pd, pr, rtol=2e-4, atol=2e-5)


def nccl_p2p_native_exchange(rank, world, device_type):
    """contrib.nccl_p2p over its own ncclComm (csrc/nccl_p2p.cpp): a real unique id broadcast from rank 0, grouped send / recv on the current
    stream, open-chain ends zeroed; compared with the values the neighbours hold."""
    from apex_b200.contrib.nccl_p2p import nccl_p2p as P
    dev = torch.device("cuda", rank)
    uid = P.get_unique_nccl_id(1).to(dev)
    dist.broadcast(uid, src=0)
    handle = P.init_nccl_comm(uid, rank, world)
    assert P._groups[handle][1] is not None, "the native communicator was not built"
    for shape, dtype in (((2, 8, 1, 16), torch.float16), ((3, 5, 7), torch.float32)):
        lo_out = torch.full(shape, float(10 * rank + 1), device=dev, dtype=dtype)
        hi_out = torch.full(shape, float(10 * rank + 2), device=dev, dtype=dtype)
        lo_in, hi_in = P.left_right_halo_exchange(handle, rank == 0, rank == world - 1, lo_out, hi_out)
        want_lo = 0.0 if rank == 0 else float(10 * (rank - 1) + 2)           # the left neighbour's right-going halo
        want_hi = 0.0 if rank == world - 1 else float(10 * (rank + 1) + 1)
        torch.testing.assert_close(lo_in, torch.full_like(lo_in, want_lo))
        torch.testing.assert_close(hi_in, torch.full_like(hi_in, want_hi))
    P.destroy_nccl_comm(handle)


def symmetric_pool_peer_map(rank, world, device_type):
    """Tensors allocated through the library's pluggable allocator (torch.cuda.MemPool) are peer-mapped after the fact: every rank reads the
    others' tensors through the returned addresses (NVLink P2P loads issued by an ordinary torch copy from the aliased pointer)."""
    from apex_b200.contrib.nccl_allocator import nccl_allocator as A
    from apex_b200.parallel.symmetric import _tensor_from_ptr
    dev = torch.device("cuda", rank)
    with A.symmetric_mem():
        _pad = torch.empty(1000, device=dev)                       # noqa: F841  (an unrelated allocation in front)
        t = torch.full((3, 257), float(rank + 1), device=dev)
    torch.cuda.synchronize()
    ptrs = A.peer_map(t, None)
    assert len(ptrs) == world and ptrs[rank] == t.data_ptr()
    dist.barrier()
    for r in range(world):
        alias = _tensor_from_ptr(ptrs[r], t.numel() * 4, dev, None).view(torch.float32).view(3, 257)
        torch.testing.assert_close(alias.clone(), torch.full((3, 257), float(r + 1), device=dev))
    dist.barrier()
    t.mul_(2)                                                      # writes are visible through the peers' mappings
    torch.cuda.synchronize()
    dist.barrier()
    alias = _tensor_from_ptr(ptrs[(rank + 1) % world], t.numel() * 4, dev, None).view(torch.float32)
    assert float(alias[0]) == 2.0 * ((rank + 1) % world + 1)
    dist.barrier()


def peer_halo_exchange_matches_allgather(rank, world, device_type):
    """PeerHaloExchanger1d (one fused P2P kernel) against halos built from an all-gather of the interiors; NCHW, channels-last and
    explicit NHWC, H- and W-split, fp16 and fp32, repeated to exercise the parity double-buffering."""
    from apex_b200.contrib.peer_memory import PeerHaloExchanger1d, PeerMemoryPool
    dev = torch.device("cuda", rank)
    pool = PeerMemoryPool(64 << 20, 8 << 20, list(range(world)))
    hh = 2
    for dtype in (torch.float16, torch.float32):
        for layout in ("nchw", "cl", "nhwc"):
            for H_split in (True, False):
                ex = PeerHaloExchanger1d(list(range(world)), rank, pool, hh)
                for it in range(3):
                    torch.manual_seed(1000 * it + rank)
                    N, C, H, W = 2, 16, 12, 8
                    explicit = layout == "nhwc"
                    if explicit:
                        shape = [N, H + 2 * hh, W, C] if H_split else [N, H, W + 2 * hh, C]
                        y = torch.randn(shape, device=dev).to(dtype)
                        dim = 1 if H_split else 2
                    else:
                        shape = [N, C, H + 2 * hh, W] if H_split else [N, C, H, W + 2 * hh]
                        y = torch.randn(shape, device=dev).to(dtype)
                        if layout == "cl":
                            y = y.contiguous(memory_format=torch.channels_last)
                        dim = 2 if H_split else 3
                    L = y.shape[dim] - 2 * hh
                    interior = y.narrow(dim, hh, L).contiguous()
                    parts = [torch.empty_like(interior) for _ in range(world)]
                    dist.all_gather(parts, interior)
                    want = y.clone()
                    lo = parts[rank - 1].narrow(dim, L - hh, hh) if rank > 0 else torch.zeros_like(interior.narrow(dim, 0, hh))
                    hi = parts[rank + 1].narrow(dim, 0, hh) if rank < world - 1 else torch.zeros_like(interior.narrow(dim, 0, hh))
                    want.narrow(dim, 0, hh).copy_(lo)
                    want.narrow(dim, L + hh, hh).copy_(hi)
                    ex(y, H_split=H_split, explicit_nhwc=explicit)
                    torch.testing.assert_close(y, want, rtol=0, atol=0)


def nvls_allreduce_matches_nccl(rank, world, device_type):
    """EXPERIMENTAL one-shot NVSwitch all-reduce (csrc/experimental/nvls_allreduce.cu) against NCCL: fp32 / bf16 / fp16, sizes that are and
    are not multiples of the per-rank vector unit, a post-scale, repeated calls on the same staging buffer, and through DDP."""
    from apex_b200.parallel.nvls_allreduce import NvlsAllReduce
    dev = torch.device("cuda", rank)
    ar = NvlsAllReduce(dist.group.WORLD, dev, 64 << 20)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2), (torch.float16, 2e-3)):
        for n in (8, 1000, 1 << 20, (1 << 22) + 13):
            for it in range(2):
                torch.manual_seed(100 * it + rank)
                x = torch.randn(n, device=dev).to(dtype)
                want = x.float().clone()
                dist.all_reduce(want)
                got = ar.allreduce_(x.clone(), scale=0.5 if it else 1.0)
                torch.testing.assert_close(got.float(), want * (0.5 if it else 1.0), rtol=tol, atol=tol * world)


def syncbn_generic_matches_concatenated_batchnorm(rank, world, device_type, uneven, fuse_relu):
    """The torch.distributed implementation of SyncBatchNorm (multi-node groups, CPU / gloo): statistics over uneven per-rank batches,
    gradients, running statistics and the fused residual-add + ReLU variant against BatchNorm over the concatenated batch."""
    from apex_b200.parallel import SyncBatchNorm
    torch.manual_seed(0)
    C = 6
    nb = [4, 7, 2, 5][:world] if uneven else [5] * world
    full = torch.randn(sum(nb), C, 5, 3) * 1.5 + 0.7
    res_full = torch.randn(sum(nb), C, 5, 3)
    dy_full = torch.randn(sum(nb), C, 5, 3)
    lo = sum(nb[:rank])
    sl = slice(lo, lo + nb[rank])
    x = full[sl].clone().requires_grad_(True)
    z = res_full[sl].clone().requires_grad_(True) if fuse_relu else None
    sbn = SyncBatchNorm(C, fuse_relu=fuse_relu)
    bn = torch.nn.BatchNorm2d(C)
    xr, zr = full.clone().requires_grad_(True), res_full.clone().requires_grad_(True)
    for _ in range(2):
        x.grad = None
        y = sbn(x, z) if fuse_relu else sbn(x)
        y.backward(dy_full[sl])
        xr.grad = None
        yr = torch.relu(bn(xr) + zr) if fuse_relu else bn(xr)
        yr.backward(dy_full)
    torch.testing.assert_close(y, yr[sl], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(x.grad, xr.grad[sl], atol=1e-5, rtol=1e-4)
    if fuse_relu:
        torch.testing.assert_close(z.grad, zr.grad[sl], atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(sbn.running_mean, bn.running_mean, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(sbn.running_var, bn.running_var, atol=1e-5, rtol=1e-4)
    gw, gb = sbn.weight.grad.clone(), sbn.bias.grad.clone()
    dist.all_reduce(gw)
    dist.all_reduce(gb)
    torch.testing.assert_close(gw, bn.weight.grad, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(gb, bn.bias.grad, atol=1e-4, rtol=1e-4)


def halo_exchangers_match_slices_of_the_full_tensor(rank, world, device_type):
    """AllGather and SendRecv halo exchangers (+ HaloPadder) on an H-split tensor: every rank must receive exactly the rows its neighbours
    own next to the cut; the outer edges of the first / last rank are zero (reference test_peer_halo_exchange_module.py:284-335 checks the
    same property for the peer-memory exchanger)."""
    from apex_b200.contrib.bottleneck.halo_exchangers import HaloExchangerAllGather, HaloExchangerSendRecv, HaloPadder
    torch.manual_seed(0)
    N, C, Hs, W, hh = 2, 3, 6, 5, 2
    full = torch.randn(N, C, Hs * world, W)
    mine = full[:, :, rank * Hs:(rank + 1) * Hs].contiguous()
    want = torch.zeros(N, C, Hs + 2 * hh, W)
    want[:, :, hh:hh + Hs] = mine
    if rank > 0:
        want[:, :, :hh] = full[:, :, rank * Hs - hh:rank * Hs]
    if rank < world - 1:
        want[:, :, hh + Hs:] = full[:, :, (rank + 1) * Hs:(rank + 1) * Hs + hh]
    ranks = list(range(world))
    for ex in (HaloExchangerAllGather(ranks, rank, dist.group.WORLD), HaloExchangerSendRecv(ranks, rank)):
        for explicit_nhwc in (False, True):
            y = mine.permute(0, 2, 3, 1).contiguous() if explicit_nhwc else mine
            got = HaloPadder(ex)(y, hh, explicit_nhwc, True)
            got = got.permute(0, 3, 1, 2) if explicit_nhwc else got
            torch.testing.assert_close(got, want, rtol=0, atol=0)
        # the in-place form writes into caller-provided halos
        li, ri = torch.full((N, C, hh, W), 7.0), torch.full((N, C, hh, W), 7.0)
        ex.left_right_halo_exchange(mine[:, :, :hh].contiguous(), mine[:, :, -hh:].contiguous(), li, ri)
        torch.testing.assert_close(li, want[:, :, :hh], rtol=0, atol=0)
        torch.testing.assert_close(ri, want[:, :, hh + Hs:], rtol=0, atol=0)


def group_batchnorm_spans_only_its_group(rank, world, device_type):
    """groupbn.BatchNorm2d_NHWC(bn_group=2) and cudnn_gbn.GroupBatchNorm2d(group_size=2) on four ranks: statistics must span the two ranks
    of a group and nothing else (reference test_cudnn_gbn_with_two_gpus.py:93-169 compares against BN over the group's concatenated batch)."""
    from apex_b200.contrib.cudnn_gbn import GroupBatchNorm2d
    from apex_b200.contrib.groupbn import BatchNorm2d_NHWC
    torch.manual_seed(0)
    C, per = 5, 3
    full = torch.randn(world * per, C, 4, 3) * (1.0 + torch.arange(world * per).view(-1, 1, 1, 1))   # every sample has its own scale
    grp = rank // 2
    mine = full[rank * per:(rank + 1) * per]
    ref = torch.nn.functional.batch_norm(full[grp * 2 * per:(grp + 1) * 2 * per], None, None, training=True)[(rank % 2) * per:(rank % 2 + 1) * per]
    nhwc = BatchNorm2d_NHWC(C, bn_group=2)
    got = nhwc(mine.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
    torch.testing.assert_close(got, ref, atol=1e-5, rtol=1e-5)
    gbn = GroupBatchNorm2d(C, group_size=2)
    got = gbn(mine.contiguous(memory_format=torch.channels_last))
    torch.testing.assert_close(got, ref, atol=1e-5, rtol=1e-5)
    # raw bnp entry points over the same group of two (forward, then the backward against autograd through BN of the group's batch)
    from apex_b200.contrib.groupbn import raw_ext as bnp
    both = full[grp * 2 * per:(grp + 1) * 2 * per].clone().requires_grad_(True)
    w, b = torch.rand(C) + 0.5, torch.randn(C)
    yref = torch.nn.functional.batch_norm(both, None, None, w, b, training=True)
    gy = torch.randn(world * per, C, 4, 3)[grp * 2 * per:(grp + 1) * 2 * per]
    dref = torch.autograd.grad(yref, both, gy)[0][(rank % 2) * per:(rank % 2 + 1) * per]
    mm, mi, rm, rv = torch.empty(C), torch.empty(C), torch.zeros(C), torch.ones(C)
    tail = (None, None, None, None, 2, torch.IntTensor([2]), 2, 100, False)
    xn = mine.permute(0, 2, 3, 1).contiguous()
    y = bnp.bn_fwd_nhwc(xn, w, b, rm, rv, mm, mi, None, 0.1, 1e-5, False, *tail)
    torch.testing.assert_close(y.permute(0, 3, 1, 2), yref[(rank % 2) * per:(rank % 2 + 1) * per].detach(), atol=1e-5, rtol=1e-5)
    dx, _, _ = bnp.bn_bwd_nhwc(xn, gy[(rank % 2) * per:(rank % 2 + 1) * per].permute(0, 2, 3, 1).contiguous(), w, b, rm, rv, mm, mi, None, 0.1,
                               1e-5, False, *tail)
    torch.testing.assert_close(dx.permute(0, 3, 1, 2), dref, atol=1e-5, rtol=1e-4)


def one_rank_raises_while_the_others_wait(rank, world, device_type):
    """Harness behaviour under failure: rank 1 dies immediately, the others block in a collective."""
    if rank == 1:
        raise ValueError("deliberate failure on rank 1")
    dist.barrier()


from apex_b200.distributed_testing.distributed_test_base import GlooDistributedTestBase, distributed  # noqa: E402


class GlooAllReduceCase(GlooDistributedTestBase):
    """Used by tests/test_cpu_utils.py: every rank runs the decorated body with self.rank / self.world_size set."""

    @distributed
    def test_all_reduce(self):
        t = torch.full((4,), float(self.rank + 1))
        dist.all_reduce(t)
        assert torch.equal(t, torch.full((4,), float(sum(range(1, self.world_size + 1))))), t


def ddp_matches_manual_allreduce(rank, world, device_type, delay, message_size, predivide):
    """apex_b200.parallel.DistributedDataParallel: bucketed / delayed all-reduce == averaging the per-rank grads by hand."""
    from apex_b200.parallel import DistributedDataParallel
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 13), torch.nn.Tanh(), torch.nn.Linear(13, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3)).to(dev)
    ref = copy.deepcopy(net)
    ddp = DistributedDataParallel(net, message_size=message_size, delay_allreduce=delay, gradient_predivide_factor=predivide,
                                  num_allreduce_streams=2)
    g = torch.Generator().manual_seed(10 + rank)
    for it in range(3):
        x = torch.randn(4, 7, generator=g).to(dev)
        for m in (ddp, ref):
            m.zero_grad()
        ddp(x).pow(2).sum().backward()
        ref(x).pow(2).sum().backward()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        for p, q in zip(net.parameters(), ref.parameters()):
            want = q.grad.clone()
            dist.all_reduce(want)
            torch.testing.assert_close(p.grad, want / world, rtol=1e-5, atol=1e-6)


def ddp_option_matrix(rank, world, device_type):
    """DistributedDataParallel options of the reference API: allreduce_always_fp32 + retain_allreduce_buffers on a bf16 model,
    gradient_average=False (sums), allreduce_trigger_params (flush exactly at the named parameters), disable / enable_allreduce."""
    from apex_b200.parallel import DistributedDataParallel

    def nets(dtype=torch.float32):
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 4)).to(dtype)
        return net, copy.deepcopy(net)

    def check(net, ref, ddp, x, scale, tol):
        for m in (net, ref):
            m.zero_grad()
        ddp(x).float().pow(2).sum().backward()
        ref(x).float().pow(2).sum().backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            want = q.grad.float().clone()
            dist.all_reduce(want)
            torch.testing.assert_close(p.grad.float(), want * scale, rtol=tol, atol=tol)

    g = torch.Generator().manual_seed(10 + rank)
    # bf16 gradients reduced in fp32, the flat buffers kept for inspection
    net, ref = nets(torch.bfloat16)
    ddp = DistributedDataParallel(net, allreduce_always_fp32=True, retain_allreduce_buffers=True, message_size=40)
    check(net, ref, ddp, torch.randn(4, 6, generator=g).bfloat16(), 1.0 / world, 2e-2)
    assert ddp.allreduce_buffers and all(b.dtype == torch.float32 for b in ddp.allreduce_buffers)
    assert sum(b.numel() for b in ddp.allreduce_buffers) == sum(p.numel() for p in net.parameters())
    # sums instead of averages
    net, ref = nets()
    check(net, ref, DistributedDataParallel(net, gradient_average=False), torch.randn(4, 6, generator=g), 1.0, 1e-5)
    # buckets close exactly when a trigger parameter's gradient arrives
    net, ref = nets()
    ddp = DistributedDataParallel(net, allreduce_trigger_params=[net[2].weight, net[0].weight], retain_allreduce_buffers=True)
    check(net, ref, ddp, torch.randn(4, 6, generator=g), 1.0 / world, 1e-5)
    assert len(ddp.allreduce_buffers) >= 2
    # disable_allreduce: gradients stay local until it is enabled again
    net, ref = nets()
    ddp = DistributedDataParallel(net)
    ddp.disable_allreduce()
    x = torch.randn(4, 6, generator=g)
    ddp(x).pow(2).sum().backward()
    ref(x).pow(2).sum().backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        torch.testing.assert_close(p.grad, q.grad)
    ddp.enable_allreduce()
    check(net, ref, ddp, torch.randn(4, 6, generator=g), 1.0 / world, 1e-5)


def ddp_reduces_when_some_parameters_get_no_gradient(rank, world, device_type, delay):
    """A parameter that does not take part in the backward never fires its hook; the buckets that did fill must still be all-reduced
    (end-of-backward callback), for both the overlapped and the delayed mode, and on consecutive iterations."""
    from apex_b200.parallel import DistributedDataParallel
    dev = torch.device(device_type, rank) if device_type == "cuda" else torch.device("cpu")

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.used, self.unused = torch.nn.Linear(8, 8), torch.nn.Linear(8, 8)

        def forward(self, x, both):
            y = self.used(x)
            return y + self.unused(x) if both else y

    torch.manual_seed(0)
    net = Net().to(dev)
    ddp = DistributedDataParallel(net, delay_allreduce=delay, message_size=1 << 30)
    for it, both in enumerate((False, True, False)):
        net.zero_grad(set_to_none=True)
        x = torch.full((4, 8), float(rank + 1 + it), device=dev)
        ddp(x, both).sum().backward()
        mean_input = sum(float(r + 1 + it) for r in range(world)) / world
        want = torch.full((8, 8), 4.0 * mean_input, device=dev)       # d(sum)/dW = sum over the batch of x, averaged over ranks
        torch.testing.assert_close(net.used.weight.grad, want)
        if both:
            torch.testing.assert_close(net.unused.weight.grad, want)
        else:
            assert net.unused.weight.grad is None


def reducer_averages_gradients_and_broadcasts_parameters(rank, world, device_type):
    """``Reducer``: construction broadcasts rank 0's parameters, ``reduce()`` averages the gradients (module form and tensor-list form,
    mixed dtypes through one flat buffer per dtype)."""
    from apex_b200.parallel import Reducer
    torch.manual_seed(rank)                                  # different initial weights on purpose
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    red = Reducer(net)
    for p in net.parameters():
        t = p.detach().clone()
        dist.broadcast(t, src=0)
        torch.testing.assert_close(p.detach(), t, rtol=0, atol=0)
    net(torch.full((2, 4), float(rank + 1))).sum().backward()
    mine = [p.grad.clone() for p in net.parameters()]
    red.reduce()
    for p, g in zip(net.parameters(), mine):
        dist.all_reduce(g)
        torch.testing.assert_close(p.grad, g / world)
    tensors = [torch.full((3,), float(rank)), torch.full((2, 2), float(rank), dtype=torch.float64)]
    Reducer(tensors).reduce()
    for t in tensors:
        torch.testing.assert_close(t, torch.full_like(t, sum(range(world)) / world))


def ddp_race_condition(rank, world, device_type):
    """Race detector by construction (reference tests/distributed/DDP/ddp_race_condition_test.py:27-78): two large parameters,
    message_size=1 (a bucket per parameter), several all-reduce streams, gradients with a closed form checked every iteration."""
    from apex_b200.parallel import DistributedDataParallel
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    n = 1 << (22 if device_type == "cuda" else 16)

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Parameter(torch.ones(n, device=dev))
            self.b = torch.nn.Parameter(torch.ones(n, device=dev))

        def forward(self, inp):
            return ((self.a * inp).sum() + (self.b * inp * 2).sum())

    model = DistributedDataParallel(Model(), message_size=1, num_allreduce_streams=3)
    x = torch.ones(n, device=dev)
    for it in range(6):
        model.zero_grad()
        inp = x * (rank + 1 + it)
        model(inp).backward()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        mean_scale = sum(r + 1 + it for r in range(world)) / world
        # every element must equal the analytic mean exactly (a race shows up as a stale or doubly-reduced bucket)
        assert float(model.module.a.grad.min()) == mean_scale == float(model.module.a.grad.max()), (it, float(model.module.a.grad.min()))
        assert float(model.module.b.grad.min()) == 2 * mean_scale == float(model.module.b.grad.max())


def spatial_bottleneck_matches_full(rank, world, device_type):
    """SpatialBottleneck (H split over the group + one-row halo exchange around the 3x3 conv) == Bottleneck on the whole image
    (reference apex/contrib/test/bottleneck/test_bottleneck_module.py:283-345)."""
    from apex_b200.contrib.bottleneck import Bottleneck, SpatialBottleneck
    from apex_b200.contrib.bottleneck.halo_exchangers import HaloExchangerAllGather, HaloExchangerSendRecv
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    torch.manual_seed(0)
    full = Bottleneck(16, 8, 32).to(dev)
    for bn in (full.bn1, full.bn2, full.bn3, full.downsample[1]):
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5)
    # cuDNN picks TF32 / different algorithms for the full image and for the shards: compare at TF32 resolution on CUDA
    tol = 1e-5 if device_type == "cpu" else 1e-2
    x = torch.randn(2, 16, 8 * world, 6, device=dev, requires_grad=True)
    want = full(x)
    gout = torch.randn_like(want)
    want.backward(gout)
    rows = slice(8 * rank, 8 * (rank + 1))
    for make in (lambda: HaloExchangerAllGather(list(range(world)), rank, None), lambda: HaloExchangerSendRecv(list(range(world)), rank)):
        halo_ex = make()
        sp = SpatialBottleneck(16, 8, 32, spatial_parallel_args=(world, rank, None, halo_ex, 1)).to(dev)
        sp.load_state_dict(full.state_dict())
        xs = x.detach()[:, :, rows].contiguous().requires_grad_()
        got = sp(xs)
        torch.testing.assert_close(got, want.detach()[:, :, rows], atol=tol, rtol=tol)
        # backward: the halo-row gradients travel back to the neighbours; weight gradients are partial sums over the H-shards
        got.backward(gout[:, :, rows].contiguous())
        torch.testing.assert_close(xs.grad, x.grad[:, :, rows], atol=10 * tol, rtol=10 * tol)
        for (n, p), (_, q) in zip(sp.named_parameters(), full.named_parameters()):
            g = p.grad.clone()
            dist.all_reduce(g)
            torch.testing.assert_close(g, q.grad, atol=10 * tol, rtol=10 * tol, msg=lambda m, n=n: f"{n}: {m}")


def dist_adam_state_dict_v1_round_trip(rank, world, device_type):
    """Deprecated v1 format (reference distributed_fused_adam.py:2907-3057): per-rank shards serialised to bytes and gathered on the root;
    loading hands every rank its own shard back. Continue the original and a fresh optimizer on the same data: identical parameters."""
    import copy
    import warnings
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    model = _model(dev)
    opt = DistributedFusedAdam(model.parameters(), lr=1e-2, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20, weight_decay=0.01,
                               fused_collectives=device_type == "cuda")
    g = torch.Generator().manual_seed(11)

    def one_step(m, o):
        o.zero_grad()
        x = torch.randn(5, 7, generator=g).to(dev)
        m(x).pow(2).mean().backward()
        o.step()

    for _ in range(2):
        one_step(model, opt)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sd = opt.state_dict(state_dict_format=1)
    assert (sd is None) == (rank != 0)
    if rank == 0:
        assert sd["format"] == 1 and len(sd["gathered_states"]) == world and sd["gathered_states"][1].dtype == torch.uint8
    model2 = copy.deepcopy(model)
    opt2 = DistributedFusedAdam(model2.parameters(), lr=1e-2, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20, weight_decay=0.01,
                                fused_collectives=device_type == "cuda", process_group=dist.new_group(list(range(world))))
    opt2.load_state_dict(sd)
    state = g.get_state()
    one_step(model, opt)
    g.set_state(state)
    one_step(model2, opt2)
    for a, b in zip(model.parameters(), model2.parameters()):
        assert torch.equal(a, b)
    assert opt2._global_step() == 3


def dist_lamb_e5m2_allgather(rank, world, device_type):
    """``e5m2_allgather=True`` (reference distributed_fused_lamb.py + multi_tensor_distopt_lamb_kernel.cu:276-357): the updated parameters cross
    the wire as E5M2 bytes, so every rank holds the same E5M2-representable values while the fp32 master shards keep full precision and the
    trajectory of the masters equals the plain run's first step."""
    from apex_b200.contrib.optimizers import DistributedFusedLAMB
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    torch.manual_seed(3)
    models = [_model(dev) for _ in range(2)]
    models[1].load_state_dict(models[0].state_dict())
    opts = [DistributedFusedLAMB(m.parameters(), lr=1e-2, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20, fused_collectives=False,
                                 e5m2_allgather=flag, process_group=dist.new_group(list(range(world)))) for m, flag in zip(models, (True, False))]
    g = torch.Generator().manual_seed(5)
    x = torch.randn(5, 7, generator=g).to(dev)
    for m, o in zip(models, opts):
        o.zero_grad()
        m(x).pow(2).mean().backward()
        o.step()
    for pe, pf in zip(models[0].parameters(), models[1].parameters()):
        # every value is exactly E5M2-representable and is the rounding of the full-precision result
        torch.testing.assert_close(pe.detach().float().to(torch.float8_e5m2).float(), pe.detach().float(), atol=0, rtol=0)
        torch.testing.assert_close(pe.detach().float(), pf.detach().float().to(torch.float8_e5m2).float(), atol=0, rtol=0)
    flat = torch.cat([p.detach().float().reshape(-1) for p in models[0].parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], t) for t in gathered)
    for se, sf in zip(opts[0]._segments, opts[1]._segments):
        torch.testing.assert_close(se.master, sf.master)      # the masters never see the quantisation


def cudnn_gbn_lib_group_of_two(rank, world, device_type):
    """``cudnn_gbn_lib.forward / backward`` with group_size = 2: every rank holds half of the batch; outputs, running statistics and gradients
    must equal plain batch norm over the concatenated batch (reference apex/contrib/cudnn_gbn/batch_norm.py:34-69)."""
    from apex_b200 import ext_compat as E
    g = E.extension_modules()["cudnn_gbn_lib"]
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    torch.manual_seed(0)
    full = torch.randn(2 * world, 8, 5, 5)
    dy_full = torch.randn(2 * world, 8, 5, 5)
    w, b = torch.randn(8), torch.randn(8)
    x = full[2 * rank:2 * rank + 2].contiguous(memory_format=torch.channels_last).to(dev)
    dy = dy_full[2 * rank:2 * rank + 2].contiguous(memory_format=torch.channels_last).to(dev)
    rm, rv = torch.zeros(8, device=dev), torch.ones(8, device=dev)
    mm, miv = torch.empty(8, device=dev), torch.empty(8, device=dev)
    y = g.forward(x, w.to(dev), b.to(dev), rm, rv, mm, miv, 0.1, 1e-5, world, rank, [])
    rm2, rv2 = torch.zeros(8), torch.ones(8)
    fr, wr, br = full.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.batch_norm(fr, rm2, rv2, wr, br, True, 0.1, 1e-5)
    ref.backward(dy_full)
    torch.testing.assert_close(y.cpu(), ref.detach()[2 * rank:2 * rank + 2], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(rm.cpu(), rm2, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(rv.cpu(), rv2, atol=1e-6, rtol=1e-5)
    dx, dw, db = g.backward(x, dy, w.to(dev), mm, miv, 1e-5, world, rank, [])
    torch.testing.assert_close(dx.cpu(), fr.grad[2 * rank:2 * rank + 2], atol=1e-5, rtol=1e-4)
    # dscale / dbias are this rank's contribution (the reference leaves their reduction to DDP)
    tot_w, tot_b = dw.clone(), db.clone()
    dist.all_reduce(tot_w)
    dist.all_reduce(tot_b)
    torch.testing.assert_close(tot_w.cpu(), wr.grad, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(tot_b.cpu(), br.grad, atol=1e-4, rtol=1e-4)


def dist_adam_fragments_partition_params(rank, world, device_type):
    """``param_fragments``: across the ranks the local-shard sub-ranges cover every element of every parameter exactly once, and the
    master shard holds the parameter's values at ``shard_range`` of the bucket's local shard."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    torch.manual_seed(3)
    params = [torch.nn.Parameter(torch.randn(n, device=dev)) for n in (9000, 130, 4097)]
    opt = DistributedFusedAdam(params, lr=1e-2, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20, fused_collectives=False)
    cover = [torch.zeros(p.numel(), dtype=torch.int32) for p in params]
    for p, c in zip(params, cover):
        for f in opt.param_fragments(p):
            if f.in_local_shard:
                c[f.shard_param_range[0]:f.shard_param_range[1]] += 1
                n = f.shard_param_range[1] - f.shard_param_range[0]
                assert f.shard_range[1] - f.shard_range[0] == n == f.shard_bucket_range[1] - f.shard_bucket_range[0]
                assert f.bucket_range[0] <= f.shard_bucket_range[0] and f.shard_bucket_range[1] <= f.bucket_range[1]
                seg = opt._segments[0]
                lo = (f.bucket_id * seg.shard_elems) + f.shard_range[0]
                torch.testing.assert_close(seg.master[lo:lo + n].cpu(), p.detach().flatten()[f.shard_param_range[0]:f.shard_param_range[1]].cpu())
            else:
                assert f.shard_range is None and f.shard_param_range is None
    for c in cover:
        dist.all_reduce(c)
        assert bool((c == 1).all())


def spatial_bottleneck_function_matches_full(rank, world, device_type):
    """SpatialBottleneckFunction.apply with the reference's argument list == the module on the whole image, on this rank's rows."""
    from apex_b200.contrib.bottleneck import Bottleneck
    from apex_b200.contrib.bottleneck.bottleneck import SpatialBottleneckFunction
    from apex_b200.contrib.bottleneck.halo_exchangers import HaloExchangerSendRecv
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    torch.manual_seed(0)
    full = Bottleneck(16, 8, 32).to(dev)
    norms = [full.bn1, full.bn2, full.bn3, full.downsample[1]]
    for bn in norms:
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5)
    x = torch.randn(2, 16, 6 * world, 5, device=dev)
    want = full(x)
    rows = slice(6 * rank, 6 * (rank + 1))
    scale, bias = zip(*(bn.get_scale_bias() for bn in norms))
    convs = [full.conv1.weight, full.conv2.weight, full.conv3.weight, full.downsample[0].weight]
    got = SpatialBottleneckFunction.apply(world, rank, None, HaloExchangerSendRecv(list(range(world)), rank), 1, False, False, (1, 1),
                                          list(scale), list(bias), None, None, x[:, :, rows].contiguous(), *convs)
    tol = 1e-5 if device_type == "cpu" else 1e-2
    torch.testing.assert_close(got, want[:, :, rows], atol=tol, rtol=tol)


def permutation_sync_uses_rank0(rank, world, device_type):
    """sync_permutations: whatever each rank found, all ranks apply rank 0's permutation (weights differ per rank here, so the local
    searches disagree), and the function of the network is preserved on every rank."""
    from apex_b200.contrib.sparsity.permutation_lib import Permutation
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    torch.manual_seed(100 + rank)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8)).to(dev)
    x = torch.randn(4, 16, device=dev)
    want = net(x)
    roots, ok = Permutation.build_fx_graph(net)
    assert ok
    Permutation.find_permutations(roots)
    local = [s.permutation for s in roots if s.permutation is not None]
    Permutation.sync_permutations(roots)
    synced = [s.permutation for s in roots if s.permutation is not None]
    gathered = [None] * world
    dist.all_gather_object(gathered, (local, synced))
    assert all(g[1] == gathered[0][0] for g in gathered), "every rank must end up with rank 0's permutations"
    assert gathered[0][0] != gathered[1][0], "the test is vacuous if the local searches agree"
    Permutation.apply_permutations(roots)
    torch.testing.assert_close(net(x), want, atol=1e-5, rtol=1e-5)
