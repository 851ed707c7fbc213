"""A/B of the persistent multi-tensor grid size (APEX_B200_MT_GRID_MULT CTAs per SM): FusedAdam / FusedSGD over 10k fp32 tensors (1.0 G elements)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
from apex_b200.optimizers import FusedAdam, FusedSGD
from apex_b200.utils.timing import time_fn
g = torch.Generator().manual_seed(0)
sizes = torch.randint(1000, 200000, (10000,), generator=g).tolist()
out = {"mult": int(os.environ.get("APEX_B200_MT_GRID_MULT", 3)), "chunk": int(os.environ.get("APEX_B200_MT_CHUNK", 65536))}
for name, cls, kw, bpe in (("adam", FusedAdam, {}, 28), ("sgd", FusedSGD, {"momentum": 0.9}, 20)):
    ps = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in sizes]
    for p in ps:
        p.grad = torch.randn_like(p)
    opt = cls(ps, lr=1e-3, **kw)
    med, mn = time_fn(opt.step, 3, 10, flush=True)
    out[name + "_ms"] = med
    out[name + "_GBps"] = sum(sizes) * bpe / med / 1e6
    del ps, opt
print(json.dumps(out))
''' % ROOT
combos = [(m, 65536) for m in (3, 4, 6, 8, 12)] if len(sys.argv) < 2 else [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1:]]
for mult, chunk in combos:
    env = dict(os.environ, APEX_B200_MT_GRID_MULT=str(mult), APEX_B200_MT_CHUNK=str(chunk))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    print(line[-1] if line else json.dumps({"mult": mult, "error": r.stderr[-300:]}), flush=True)
