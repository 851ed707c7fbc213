import sys; sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import random, torch, warnings
from apex_b200.contrib.optimizers import DistributedFusedLAMB
from apex_b200.optimizers import FusedLAMB
warnings.simplefilter("ignore")
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = random.Random(seed); torch.manual_seed(seed)
    shapes = []
    for _ in range(rng.randint(1, 6)):
        kind = rng.choice(["vec", "mat", "big", "odd"])
        shapes.append({"vec": (rng.randint(1, 300),), "mat": (rng.randint(1, 40), rng.randint(1, 40)), "big": (rng.randint(1000, 9000),),
                       "odd": (rng.randint(1, 7), rng.randint(1, 7), rng.randint(1, 7))}[kind])
    ps = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    common = dict(lr=1e-2, eps=1e-6, weight_decay=rng.choice([0.0, 0.01]), max_grad_norm=rng.choice([0.0, 1.0]), use_nvlamb=rng.random() < 0.5,
                  adam_w_mode=rng.random() < 0.7, grad_averaging=rng.random() < 0.7, bias_correction=rng.random() < 0.8)
    try:
        opt = DistributedFusedLAMB(ps, device="cpu", bucket_cap_mb=rng.choice([0.0005, 0.004, 1.0]), **common)
        ref = FusedLAMB(qs, **common)
        for step in range(3):
            if step: opt.zero_grad()
            for p, q in zip(ps, qs):
                g = torch.randn(p.shape); q.grad = g.clone(); p.grad = g.clone()
            opt.step(); ref.step()
        for p, q in zip(ps, qs):
            if not torch.allclose(p, q, atol=2e-5, rtol=2e-5):
                raise AssertionError(f"mismatch {tuple(p.shape)} {(p-q).abs().max().item()}")
    except Exception as e:
        print("seed", seed, type(e).__name__, str(e)[:200], common); bad += 1
print("bad", bad)
