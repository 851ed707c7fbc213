"""A/B of the single-GPU ZeRO step instantiations (APEX_B200_DIST_W1 = 0..5, grid = 148 x k): ms for 2^30 bf16 parameters.
Each variant runs in its own process (the knob is read once)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
from apex_b200.contrib.optimizers import DistributedFusedAdam
from apex_b200.utils.timing import time_fn
dev = torch.device("cuda:0")
ps = [torch.nn.Parameter(torch.randn(1 << 30, device=dev, dtype=torch.bfloat16))]
opt = DistributedFusedAdam(ps, lr=1e-3, weight_decay=0.1, capturable=True)
opt.zero_grad(); ps[0].grad.normal_()
import apex_b200.contrib.optimizers.distributed_fused_adam as M
grid = int(os.environ.get("GRID_MULT", "3")) * 148
orig = opt._launch
def launch(seg, mode, group, step, *a, **k):
    k.setdefault("grid", grid)
    return orig(seg, mode, group, step, *a, **k)
opt._launch = launch
med, mn = time_fn(opt.step, 3, 10, flush=False)
print(json.dumps({"variant": int(os.environ.get("APEX_B200_DIST_W1", 0)), "grid_mult": grid // 148, "ms": med, "min_ms": mn, "GBps": (1 << 30) * 28 / med / 1e6}))
''' % ROOT
for variant, mults in ((0, (4, 6, 8, 12)), (2, (3, 6))):
    for gm in mults:
        env = dict(os.environ, APEX_B200_DIST_W1=str(variant), GRID_MULT=str(gm))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        print(line[-1] if line else json.dumps({"variant": variant, "error": r.stderr[-300:]}), flush=True)
