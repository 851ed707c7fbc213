"""Single-GPU op micro-benchmarks: achieved GB/s vs the measured HBM peak (MEASURED_PEAKS.json).
Inputs are larger than L2 and L2 is flushed between iterations; CUDA-event timed."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apex_b200.utils.timing import measured_peaks, time_fn  # noqa: E402


def bench_adam(n_tensors, out):
    from apex_b200.optimizers import FusedAdam
    g = torch.Generator().manual_seed(0)
    if n_tensors == 1:
        sizes = [512 * 1024 * 1024]
    else:
        sizes = torch.randint(1000, 200000, (n_tensors,), generator=g).tolist()
    ps = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in sizes]
    for p in ps:
        p.grad = torch.randn_like(p)
    opt = FusedAdam(ps, lr=1e-3, weight_decay=0.01)
    opt.step()
    med, mn = time_fn(lambda: opt.step(), warmup=3, iters=10)
    nel = sum(sizes)
    gb = nel * 28 / 1e9
    out.append({"op": f"FusedAdam fp32 {n_tensors} tensors", "elems": nel, "ms": med, "ms_min": mn, "GBps": gb / med * 1e3})
    # torch fused reference on the same box
    topt = torch.optim.AdamW(ps, lr=1e-3, weight_decay=0.01, fused=True)
    topt.step()
    med2, _ = time_fn(lambda: topt.step(), warmup=2, iters=5)
    out.append({"op": f"torch AdamW(fused=True) {n_tensors} tensors", "elems": nel, "ms": med2, "GBps": gb / med2 * 1e3})
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        opt.step()
    t1 = time.perf_counter()
    out.append({"op": f"FusedAdam host-side enqueue time {n_tensors} tensors", "ms": (t1 - t0) / 20 * 1e3})
    torch.cuda.synchronize()


def bench_lamb(n_tensors, out):
    from apex_b200.optimizers import FusedLAMB
    g = torch.Generator().manual_seed(0)
    sizes = torch.randint(1000, 200000, (n_tensors,), generator=g).tolist()
    ps = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in sizes]
    for p in ps:
        p.grad = torch.randn_like(p)
    opt = FusedLAMB(ps, lr=1e-3)
    opt.step()
    med, mn = time_fn(lambda: opt.step(), warmup=3, iters=10)
    nel = sum(sizes)
    gb = nel * (4 + 16 + 12 + 8 + 4) / 1e9  # norm read g; pass1 r g,p,m,v w g,m,v; pass2 r g,p w p
    out.append({"op": f"FusedLAMB fp32 {n_tensors} tensors", "elems": nel, "ms": med, "ms_min": mn, "GBps": gb / med * 1e3})


def bench_norm(out):
    from apex_b200.normalization import FusedLayerNorm, FusedRMSNorm
    for hidden in (1024, 4096, 8192, 16384):
        rows = (1 << 30) // (hidden * 2)  # 1 GiB bf16 activations
        for name, M in (("LayerNorm", FusedLayerNorm), ("RMSNorm", FusedRMSNorm)):
            m = M(hidden).cuda().bfloat16()
            x = torch.randn(rows, hidden, device="cuda", dtype=torch.bfloat16, requires_grad=True)
            dy = torch.randn_like(x)
            y = m(x)
            med, mn = time_fn(lambda: m(x), warmup=3, iters=10)
            out.append({"op": f"{name} fwd bf16 h={hidden}", "rows": rows, "ms": med, "GBps": rows * hidden * 4 / 1e9 / med * 1e3})
            y = m(x)
            med, mn = time_fn(lambda: torch.autograd.grad(y, (x, m.weight), dy, retain_graph=True), warmup=3, iters=10)
            out.append({"op": f"{name} bwd bf16 h={hidden}", "rows": rows, "ms": med, "GBps": rows * hidden * 6 / 1e9 / med * 1e3})
            if name == "LayerNorm":
                tm = torch.nn.LayerNorm(hidden).cuda().bfloat16()
                med, _ = time_fn(lambda: tm(x), warmup=2, iters=5)
                out.append({"op": f"torch LayerNorm fwd bf16 h={hidden}", "ms": med, "GBps": rows * hidden * 4 / 1e9 / med * 1e3})
            del x, dy, y


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="adam,lamb,norm")
    ap.add_argument("--out", default="gpurun_out/bench_ops.json")
    a = ap.parse_args()
    res = []
    pk = measured_peaks()
    if "adam" in a.what:
        bench_adam(1, res)
        bench_adam(10000, res)
    if "lamb" in a.what:
        bench_lamb(10000, res)
    if "norm" in a.what:
        bench_norm(res)
    for r in res:
        if "GBps" in r:
            r["frac_of_hbm_peak"] = r["GBps"] / pk["hbm_gbs"]
        print(json.dumps(r))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"peaks": pk, "results": res}, open(a.out, "w"), indent=1)
