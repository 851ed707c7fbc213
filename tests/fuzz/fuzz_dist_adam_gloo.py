import sys; sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import random, torch, warnings
import torch.distributed as dist

def case(rank, world, device_type, lo, hi):
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    warnings.simplefilter("ignore")
    bad = 0
    for seed in range(lo, hi):
        rng = random.Random(seed); torch.manual_seed(seed)
        ps = []
        for _ in range(rng.randint(1, 6)):
            kind = rng.choice(["vec", "mat", "big", "scalar", "odd"])
            shape = {"vec": (rng.randint(1, 300),), "mat": (rng.randint(1, 40), rng.randint(1, 40)), "big": (rng.randint(1000, 9000),),
                     "scalar": (), "odd": (rng.randint(1, 7), rng.randint(1, 7), rng.randint(1, 7))}[kind]
            ps.append(torch.nn.Parameter(torch.randn(shape)))
        qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        kw = dict(bucket_cap_mb=rng.choice([0.0005, 0.004, 0.05, 1.0]) * world, weight_decay=rng.choice([0.0, 0.01]))
        extra = rng.choice([{}, {"store_params": True}, {"overlap_grad_sync": True}, {"overlap_grad_sync": False}])
        opt = DistributedFusedAdam(ps, lr=1e-2, device="cpu", fused_collectives=False, **kw, **extra)
        ref = torch.optim.AdamW(qs, lr=1e-2, weight_decay=kw["weight_decay"])
        try:
            for step in range(3):
                if step: opt.zero_grad(set_to_none=rng.random() < 0.5)
                grng = torch.Generator().manual_seed(seed * 100 + step)          # the same on every rank -> per-rank grads differ by rank scaling
                for p, q in zip(ps, qs):
                    base = torch.randn(p.shape, generator=grng)
                    mine = base * (rank + 1)
                    q.grad = base * sum(r + 1 for r in range(world)) / world     # what the average over ranks is
                    if p.grad is None: p.grad = mine.clone()
                    else: p.grad.copy_(mine)
                opt.step(); ref.step()
            for p, q in zip(ps, qs):
                if not torch.allclose(p, q, atol=2e-5, rtol=2e-5):
                    raise AssertionError(f"mismatch {tuple(p.shape)} {(p-q).abs().max().item()}")
            sd = opt.state_dict()
            opt.load_state_dict(sd)
        except Exception as e:
            print("rank", rank, "seed", seed, type(e).__name__, str(e)[:200], kw, extra, flush=True); bad += 1
    t = torch.tensor([bad]); dist.all_reduce(t)
    if rank == 0: print("world", world, "bad", int(t), flush=True)

if __name__ == "__main__":
    from apex_b200.testing.dist_harness import run_distributed
    world = int(sys.argv[1])
    run_distributed(case, world, "cpu", int(sys.argv[2]), int(sys.argv[3]), backend="gloo")
