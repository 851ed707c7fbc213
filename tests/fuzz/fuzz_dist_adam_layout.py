import sys; sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import random, torch, warnings
from apex_b200.contrib.optimizers import DistributedFusedAdam
warnings.simplefilter("ignore")
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = random.Random(seed); torch.manual_seed(seed)
    ngroups = rng.randint(1, 3)
    groups, refgroups = [], []
    for g in range(ngroups):
        ps = []
        for _ in range(rng.randint(1, 5)):
            kind = rng.choice(["vec", "mat", "big", "scalar", "odd"])
            shape = {"vec": (rng.randint(1, 300),), "mat": (rng.randint(1, 40), rng.randint(1, 40)), "big": (rng.randint(1000, 9000),),
                     "scalar": (), "odd": (rng.randint(1, 7), rng.randint(1, 7), rng.randint(1, 7))}[kind]
            ps.append(torch.nn.Parameter(torch.randn(shape)))
        opts = {"lr": rng.choice([1e-2, 1e-3]), "weight_decay": rng.choice([0.0, 0.01])}
        groups.append({"params": ps, **opts})
        refgroups.append({"params": [torch.nn.Parameter(p.detach().clone()) for p in ps], **opts})
    kw = dict(bucket_cap_mb=rng.choice([0.0005, 0.004, 0.05, 1.0]), adam_w_mode=rng.random() < 0.7)
    extra = rng.choice([{}, {"store_params": True}, {"overlap_grad_sync": False}, {"contiguous_grad_buffer": True}])
    try:
        opt = DistributedFusedAdam(groups, lr=1e-3, device="cpu", **kw, **extra)
    except TypeError:
        opt = DistributedFusedAdam(groups, lr=1e-3, device="cpu", **kw)
    ref = (torch.optim.AdamW if kw["adam_w_mode"] else torch.optim.Adam)(refgroups, lr=1e-3)
    try:
        for step in range(3):
            set_none = rng.random() < 0.5
            opt.zero_grad(set_to_none=set_none) if step else None
            for g, rg in zip(groups, refgroups):
                for p, q in zip(g["params"], rg["params"]):
                    if rng.random() < 0.1 and step:       # a parameter that gets no gradient this step
                        q.grad = torch.zeros_like(q); gr = torch.zeros_like(p)
                    else:
                        gr = torch.randn_like(p); q.grad = gr.clone()
                    if p.grad is None: p.grad = gr.clone()
                    else: p.grad.copy_(gr)
            opt.step(); ref.step()
        for g, rg in zip(groups, refgroups):
            for p, q in zip(g["params"], rg["params"]):
                if not torch.allclose(p, q, atol=1e-5, rtol=1e-5):
                    raise AssertionError(f"mismatch {tuple(p.shape)} {(p-q).abs().max().item()}")
        sd = opt.state_dict()
        opt.load_state_dict(sd)
    except Exception as e:
        print("seed", seed, type(e).__name__, str(e)[:200], kw, extra); bad += 1
print("bad", bad)
