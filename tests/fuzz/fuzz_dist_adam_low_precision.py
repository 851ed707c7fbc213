import sys; sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
import random, torch, warnings
from apex_b200.contrib.optimizers import DistributedFusedAdam
warnings.simplefilter("ignore")
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = random.Random(seed); torch.manual_seed(seed)
    variant = rng.choice(["bf16_rem", "bf16_master", "fp16_master", "scaled", "bf16_gradsync"])
    pdt = torch.float16 if variant == "fp16_master" else torch.bfloat16
    shapes = [ (rng.randint(1, 300),) if rng.random() < 0.5 else (rng.randint(1, 40), rng.randint(1, 40)) for _ in range(rng.randint(1, 5))]
    master = [torch.randn(s) for s in shapes]
    ps = [torch.nn.Parameter(m.to(pdt)) for m in master]
    qs = [torch.nn.Parameter(p.detach().float().clone()) for p in ps]     # fp32 reference starting from the rounded values
    kw = dict(bucket_cap_mb=rng.choice([0.0005, 0.004, 1.0]), weight_decay=0.01)
    if variant == "bf16_rem": kw.update(store_params=False, store_param_remainders=True)
    elif variant == "scaled": kw.update(with_scaled_states=True, dtype=torch.float16 if rng.random() < 0.5 else torch.bfloat16)
    elif variant == "bf16_gradsync": kw.update(grad_sync_dtype=torch.bfloat16)
    try:
        opt = DistributedFusedAdam(ps, lr=1e-2, device="cpu", **kw)
        ref = torch.optim.AdamW(qs, lr=1e-2, weight_decay=0.01)
        for step in range(4):
            if step: opt.zero_grad(set_to_none=rng.random() < 0.5)
            for p, q in zip(ps, qs):
                g = torch.randn(p.shape).to(pdt)
                q.grad = g.float()
                if p.grad is None: p.grad = g.clone()
                else: p.grad.copy_(g)
            opt.step(); ref.step()
        tol = {"scaled": 3e-2}.get(variant, 1.2e-2 if pdt == torch.bfloat16 else 2e-3)
        for p, q in zip(ps, qs):
            err = (p.float() - q).abs().max().item()
            if err > tol * max(1.0, q.abs().max().item()):
                raise AssertionError(f"mismatch {tuple(p.shape)} err {err}")
        sd = opt.state_dict(); opt.load_state_dict(sd)
    except Exception as e:
        print("seed", seed, variant, type(e).__name__, str(e)[:160], kw); bad += 1
print("bad", bad)
