"""Test bodies that run inside spawned ranks (must be importable by name from child processes)."""
import copy

import torch
import torch.distributed as dist


def _model(device, dtype=torch.float32, seed=0):
    torch.manual_seed(seed)
    # 11 small layers whose sizes straddle bucket boundaries (reference test_dist_adam.py:47 uses a 71-element bucket)
    m = torch.nn.Sequential(*[torch.nn.Linear(7, 7) for _ in range(11)]).to(device=device, dtype=dtype)
    return m


def dist_adam_matches_ddp_adamw(rank, world, device_type, fused, steps=4, clip=False, dtype=torch.float32, grad_sync_dtype=None):
    """Oracle = every rank holds the full model + torch.optim.AdamW on all-reduced (averaged) grads."""
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank) if device_type == "cuda" else torch.device("cpu")
    ref_model = _model(dev)
    dist_model = copy.deepcopy(ref_model).to(dtype)
    groups = lambda m: [{"params": [p for n, p in m.named_parameters() if "weight" in n], "lr": 3e-3},
                        {"params": [p for n, p in m.named_parameters() if "bias" in n], "lr": 1e-2, "weight_decay": 0.0}]
    ref_opt = torch.optim.AdamW(groups(ref_model), lr=3e-3, weight_decay=0.05)
    opt = DistributedFusedAdam(groups(dist_model), lr=3e-3, weight_decay=0.05, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20,
                               fused_collectives=("auto" if fused else False), grad_sync_dtype=grad_sync_dtype)
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
    assert set(sd["state"].keys()) == set(range(22))
    assert sd["state"][0]["exp_avg"].shape == (7, 7)


def dist_adam_grad_scaler_skips_on_inf(rank, world, device_type):
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    dev = torch.device("cuda", rank)
    model = _model(dev)
    opt = DistributedFusedAdam(model.parameters(), lr=1e-2, device=dev, bucket_cap_mb=2048 * 4 * world / 2 ** 20)
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
