import sys; sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import random, torch, torch.nn as nn, copy
import torch.distributed as dist

def case(rank, world, device_type, lo, hi):
    from apex_b200.parallel import DistributedDataParallel as DDP
    bad = 0
    for seed in range(lo, hi):
        rng = random.Random(seed); torch.manual_seed(seed)
        layers, width = [], 8
        for _ in range(rng.randint(1, 5)):
            w2 = rng.choice([4, 8, 16, 33])
            layers += [nn.Linear(width, w2, bias=rng.random() < 0.7), rng.choice([nn.ReLU(), nn.Tanh(), nn.Identity()])]
            if rng.random() < 0.3: layers.append(nn.LayerNorm(w2))
            width = w2
        model = nn.Sequential(*layers)
        if rng.random() < 0.3: model = model.double() if rng.random() < 0.3 else model
        ref = copy.deepcopy(model)
        kw = dict(message_size=rng.choice([1, 50, 1000, 10_000_000]), delay_allreduce=rng.random() < 0.3,
                  retain_allreduce_buffers=False, allreduce_always_fp32=rng.random() < 0.3,
                  gradient_average=rng.random() < 0.8, gradient_predivide_factor=rng.choice([1.0, 2.0]), num_allreduce_streams=rng.choice([1, 2]))
        try:
            ddp = DDP(model, **kw)
            dt = next(model.parameters()).dtype
            for it in range(3):
                x = torch.randn(4, 8, generator=torch.Generator().manual_seed(seed * 10 + it * 2 + rank)).to(dt)
                model.zero_grad(); ref.zero_grad()
                ddp(x).pow(2).mean().backward()
                ref(x).pow(2).mean().backward()
                for p in ref.parameters():
                    dist.all_reduce(p.grad)
                    if kw["gradient_average"]: p.grad /= world
                    else: p.grad /= kw["gradient_predivide_factor"]      # upstream apex: the pre-division is only undone when averaging
                for (n, p), q in zip(model.named_parameters(), ref.parameters()):
                    if not torch.allclose(p.grad, q.grad, atol=1e-5, rtol=1e-4):
                        raise AssertionError(f"it {it} {n} {(p.grad - q.grad).abs().max().item()}")
        except Exception as e:
            print("rank", rank, "seed", seed, type(e).__name__, str(e)[:200], kw, flush=True); bad += 1
    t = torch.tensor([bad]); dist.all_reduce(t)
    if rank == 0: print("world", world, "bad", int(t), flush=True)

if __name__ == "__main__":
    from apex_b200.testing.dist_harness import run_distributed
    run_distributed(case, int(sys.argv[1]), "cpu", int(sys.argv[2]), int(sys.argv[3]), backend="gloo")
