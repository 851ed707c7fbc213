"""DistributedFusedAdam on GPUs: the one-kernel step (csrc/dist_adam.cu) vs AdamW / DDP oracles.
Single-GPU cases always run; multi-GPU cases need >= 2 devices (gpurun --gpus 2)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("pdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("two_phase", [False, True])
def test_world1_fused_kernel_matches_adamw(cuda_dev, pdtype, two_phase):
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    torch.manual_seed(0)
    shapes = [(300, 77), (4099,), (64, 64), (5,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=cuda_dev).to(pdtype)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().float().clone()) for p in ps]
    a = DistributedFusedAdam(ps, lr=1e-2, weight_decay=0.1, bucket_cap_mb=0.05)
    b = torch.optim.AdamW(qs, lr=1e-2, weight_decay=0.1)
    assert a.fused_collectives
    for it in range(4):
        a.zero_grad()
        for p, q in zip(ps, qs):
            g = torch.randn(p.shape, device=cuda_dev).to(pdtype)
            p.grad.copy_(g)
            q.grad = g.float()
        if two_phase:
            n = a.clip_grad_norm(0.5)
            nr = torch.nn.utils.clip_grad_norm_(qs, 0.5)
            torch.testing.assert_close(n, nr, rtol=1e-4, atol=1e-5)
        a.step()
        b.step()
    tol = 1e-5 if pdtype == torch.float32 else 2e-2
    for p, q in zip(ps, qs):
        torch.testing.assert_close(p.float(), q, rtol=tol, atol=tol)
    assert a.kernel_launches == (8 if two_phase else 4)
    if not two_phase:
        torch.testing.assert_close(a.last_grad_norm(), torch.stack([q.grad.norm() for q in qs]).norm(), rtol=1e-4, atol=1e-5)


def test_world1_capturable_and_state_dict(cuda_dev):
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(1000, 33, device=cuda_dev, dtype=torch.bfloat16))]
    qs = [torch.nn.Parameter(ps[0].detach().float().clone())]
    a = DistributedFusedAdam(ps, lr=1e-2, capturable=True, bucket_cap_mb=0.05)
    b = torch.optim.AdamW(qs, lr=1e-2, weight_decay=0.0)
    for it in range(3):
        a.zero_grad()
        g = torch.randn_like(ps[0])
        ps[0].grad.copy_(g)
        qs[0].grad = g.float()
        a.param_groups[0]["lr"].fill_(1e-2 * (it + 1))
        b.param_groups[0]["lr"] = 1e-2 * (it + 1)
        a.step()
        b.step()
    torch.testing.assert_close(ps[0].float(), qs[0], rtol=2e-2, atol=2e-2)
    sd = a.state_dict()
    torch.testing.assert_close(sd["state"][0]["param"].to(cuda_dev), qs[0].detach(), rtol=1e-4, atol=1e-4)
    assert int(a.param_groups[0]["step"].item()) == 3


def test_generic_path_matches_fused_path(cuda_dev):
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    torch.manual_seed(0)
    base = [torch.randn(513, 31, device=cuda_dev), torch.randn(70000, device=cuda_dev)]
    res = []
    for fused in ("auto", False):
        ps = [torch.nn.Parameter(t.clone()) for t in base]
        o = DistributedFusedAdam(ps, lr=1e-2, weight_decay=0.05, fused_collectives=fused, bucket_cap_mb=0.1)
        g = torch.Generator(device=cuda_dev).manual_seed(5)
        for _ in range(3):
            o.zero_grad()
            for p in ps:
                p.grad.copy_(torch.randn(p.shape, device=cuda_dev, generator=g))
            o.step()
        res.append([p.detach().clone() for p in ps])
    for x, y in zip(*res):
        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-6)


def test_param_remainders_generic(cuda_dev):
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    torch.manual_seed(0)
    w = torch.randn(4096, 9, device=cuda_dev)
    p = torch.nn.Parameter(w.bfloat16())
    q = torch.nn.Parameter(w.bfloat16().float())
    a = DistributedFusedAdam([p], lr=1e-2, store_param_remainders=True, bucket_cap_mb=0.05)
    b = torch.optim.AdamW([q], lr=1e-2, weight_decay=0.0)
    for _ in range(3):
        a.zero_grad()
        g = torch.randn_like(p)
        p.grad.copy_(g)
        q.grad = g.float()
        a.step()
        b.step()
    # bf16 param + int16 remainder reproduces the fp32 master exactly -> the bf16 view is the rounded fp32 value
    torch.testing.assert_close(p.float(), q.detach().bfloat16().float(), rtol=0, atol=1e-2)


# ---------------------------------------------------------------- multi-GPU (NVLink P2P / NVLS in-kernel collectives)
@pytest.mark.parametrize("clip", [False, True])
def test_two_gpus_fused_matches_ddp_adamw(cuda_dev, clip):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_matches_ddp_adamw, 2, "cuda", True, 4, clip, backend="nccl")


def test_two_gpus_fused_bf16(cuda_dev):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_matches_ddp_adamw, 2, "cuda", True, 3, False, torch.bfloat16, None, backend="nccl")


def test_two_gpus_nccl_path_matches_ddp_adamw(cuda_dev):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_matches_ddp_adamw, 2, "cuda", False, 3, True, backend="nccl")


def test_two_gpus_grad_scaler(cuda_dev):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_grad_scaler_skips_on_inf, 2, "cuda", backend="nccl")


def test_two_gpus_overlap_grad_sync(cuda_dev):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_overlap_grad_sync, 2, "cuda", backend="nccl")


def test_two_gpus_overlap_param_sync_into_fused_dense(cuda_dev):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_overlap_param_sync, 2, "cuda", backend="nccl")


def test_world1_overlap_param_sync_into_fused_dense(cuda_dev):
    """Same protocol at world size 1 (flags written and acquired on one GPU; the push runs on the side stream)."""
    import copy
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    from apex_b200.fused_dense import FusedDense
    from apex_b200.ops import gemm as G
    from apex_b200.parallel.param_sync import attach_param_sync_hooks
    torch.manual_seed(0)
    ma = torch.nn.Sequential(FusedDense(512, 1024), torch.nn.GELU(), FusedDense(1024, 512), torch.nn.LayerNorm(512)).to(cuda_dev, torch.bfloat16)
    mb = copy.deepcopy(ma)
    a = DistributedFusedAdam(ma.parameters(), lr=1e-2, bucket_cap_mb=0.25, overlap_param_sync=True)
    b = DistributedFusedAdam(mb.parameters(), lr=1e-2, bucket_cap_mb=0.25)
    attach_param_sync_hooks(ma)
    guarded0 = G.stats["guarded"]
    for it in range(5):
        x = torch.randn(256, 512, device=cuda_dev, dtype=torch.bfloat16)
        a.zero_grad()
        b.zero_grad()
        ya, yb = ma(x), mb(x)
        assert torch.equal(ya, yb), it
        ya.float().pow(2).mean().backward()
        yb.float().pow(2).mean().backward()
        a.step()
        b.step()
    a.param_sync()
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert torch.equal(pa, pb)
    assert G.stats["guarded"] - guarded0 >= 4


def test_world1_step_reads_handed_over_gradients_in_place(cuda_dev):
    """World size 1 + zero_grad(set_to_none=True): gradients are neither accumulated into nor copied to the contiguous buffer; the step kernel
    follows a per-parameter source table. Odd-sized parameters (copy fallback), micro-batch accumulation and a parameter that only
    sometimes gets a gradient must match the default (buffer-view) mode bit for bit."""
    import copy
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    torch.manual_seed(0)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b = torch.nn.Linear(64, 96), torch.nn.Linear(96, 33)     # 33-element bias / 3168-element weight: numel % 8 != 0
            self.extra = torch.nn.Parameter(torch.randn(4099))

        def forward(self, x, use_extra):
            y = self.b(torch.tanh(self.a(x)))
            return y.float().pow(2).mean() + (self.extra.float().sum() * 1e-3 if use_extra else 0.0)

    ma = Net().to(cuda_dev, torch.bfloat16)
    mb = copy.deepcopy(ma)
    a = DistributedFusedAdam(ma.parameters(), lr=1e-2, bucket_cap_mb=0.02, weight_decay=0.01)
    b = DistributedFusedAdam(mb.parameters(), lr=1e-2, bucket_cap_mb=0.02, weight_decay=0.01)
    direct = 0
    for it in range(5):
        a.zero_grad(set_to_none=True)
        b.zero_grad()
        for micro in range(2):
            x = torch.randn(8, 64, device=cuda_dev, dtype=torch.bfloat16)
            ma(x, it % 2 == 0).backward()
            mb(x, it % 2 == 0).backward()
        direct += sum(t is not None for seg in a._segments for t in seg.live)
        a.step()
        b.step()
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.equal(pa, pb), it
    assert direct >= 5, "no gradient was read in place"
    torch.testing.assert_close(a.last_grad_norm(), b.last_grad_norm(), rtol=1e-5, atol=1e-7)


def test_two_gpus_step_in_backward(cuda_dev):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_step_in_backward, 2, "cuda", backend="nccl")


def test_world1_step_in_backward_matches_step_after_backward(cuda_dev):
    import copy
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    torch.manual_seed(0)
    ma = torch.nn.Sequential(*[torch.nn.Linear(64, 64) for _ in range(5)]).to(cuda_dev, torch.bfloat16)
    mb = copy.deepcopy(ma)
    a = DistributedFusedAdam(ma.parameters(), lr=1e-2, bucket_cap_mb=0.01, overlap_step_with_backward=True, capturable=True)
    b = DistributedFusedAdam(mb.parameters(), lr=1e-2, bucket_cap_mb=0.01, capturable=True)
    for it in range(4):
        a.zero_grad(set_to_none=True)
        b.zero_grad()
        x = torch.randn(16, 64, device=cuda_dev, dtype=torch.bfloat16)
        ma(x).float().pow(2).mean().backward()
        mb(x).float().pow(2).mean().backward()
        assert sum(sum(seg.bucket_stepped) for seg in a._segments) > 1
        a.step()
        b.step()
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert torch.equal(pa, pb)
    assert int(a.param_groups[0]["step"]) == 4


def test_two_gpus_cuda_graph_capture_of_the_distributed_step(cuda_dev):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_cuda_graph_replays, 2, "cuda", backend="nccl")


def test_world1_cuda_graph_capture(cuda_dev):
    """The same capture at world size 1 (always runs on the single-GPU tier)."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    torch.manual_seed(3)
    pa = [torch.nn.Parameter(torch.randn(513, 40, device=cuda_dev).bfloat16())]
    pb = [torch.nn.Parameter(pa[0].detach().clone())]
    a = DistributedFusedAdam(pa, lr=1e-2, capturable=True, bucket_cap_mb=0.05)
    b = DistributedFusedAdam(pb, lr=1e-2, capturable=True, bucket_cap_mb=0.05)
    a.zero_grad()
    b.zero_grad()
    gen = torch.Generator(device=cuda_dev).manual_seed(9)

    def new_grads():
        gr = torch.randn(pa[0].shape, device=cuda_dev, generator=gen).bfloat16()
        pa[0].grad.copy_(gr)
        pb[0].grad.copy_(gr)

    new_grads(); a.step(); b.step()   # noqa: E702
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    new_grads()
    with torch.cuda.graph(graph):
        a.step()
    b.step()
    graph.replay()
    for _ in range(10):
        new_grads()
        graph.replay()
        b.step()
    torch.cuda.synchronize()
    assert torch.equal(pa[0], pb[0]) and int(a.param_groups[0]["step"].item()) == 12


def test_four_gpus_fused(cuda_dev):
    _need(4)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_matches_ddp_adamw, 4, "cuda", True, 3, True, backend="nccl")


def test_eight_gpus_fused(cuda_dev):
    _need(8)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_adam_matches_ddp_adamw, 8, "cuda", True, 3, True, backend="nccl")


# ---------------------------------------------------------------- DistributedFusedLAMB: device flow (stage kernels + MODE_PUSH + NVLS sums)
def test_world1_dist_lamb_matches_fused_lamb_and_skips_overflow_on_the_device(cuda_dev):
    import copy
    from apex_b200.contrib.optimizers import DistributedFusedLAMB
    from apex_b200.optimizers import FusedLAMB
    torch.manual_seed(0)
    ref_model = torch.nn.Sequential(*[torch.nn.Linear(40, 40) for _ in range(4)]).to(cuda_dev)
    dist_model = copy.deepcopy(ref_model)
    ref_opt = FusedLAMB(ref_model.parameters(), lr=2e-2, weight_decay=0.01, max_grad_norm=0.5, eps=1e-6)
    opt = DistributedFusedLAMB(dist_model.parameters(), lr=2e-2, weight_decay=0.01, max_grad_norm=0.5, eps=1e-6, bucket_cap_mb=0.01)
    for it in range(4):
        opt.zero_grad()
        ref_opt.zero_grad()
        x = torch.randn(8, 40, device=cuda_dev)
        ref_model(x).pow(2).mean().backward()
        dist_model(x).pow(2).mean().backward()
        if it == 2:   # overflow step: must be skipped without reading the flag on the host, bias correction must not advance
            before = [p.detach().clone() for p in dist_model.parameters()]
            next(dist_model.parameters()).grad.view(-1)[0] = float("inf")
            opt.step()
            for p, q in zip(dist_model.parameters(), before):
                assert torch.equal(p, q)
            assert int(opt._applied_steps.item()) == 2
            continue
        ref_opt.step()
        opt.step()
        for pr, pd in zip(ref_model.parameters(), dist_model.parameters()):
            torch.testing.assert_close(pd, pr, rtol=2e-4, atol=2e-5)
    assert opt.state_dict()["state"]["step"] == 3


def test_two_gpus_dist_lamb(cuda_dev):
    _need(2)
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.dist_lamb_matches_fused_lamb, 2, "cuda", backend="nccl")
