"""LayerNorm / RMSNorm bandwidth sweep (1 GiB bf16 activations per case); prints one JSON line per case. Tuning knobs are environment
variables read by the kernels' launchers (APEX_B200_LN_*), so one process = one configuration."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apex_b200.normalization import FusedLayerNorm, FusedRMSNorm  # noqa: E402
from apex_b200.utils.timing import time_fn  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
hiddens = [int(h) for h in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024, 2048, 4096, 8192, 12288, 16384]
PEAK = 6588.7
for hidden in hiddens:
    rows = (1 << 30) // (hidden * 2)
    for name, M in (("LN", FusedLayerNorm), ("RMS", FusedRMSNorm)):
        m = M(hidden).cuda().bfloat16()
        x = torch.randn(rows, hidden, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        dy = torch.randn_like(x)
        y = m(x)
        f, _ = time_fn(lambda: m(x), warmup=3, iters=10)
        b, _ = time_fn(lambda: torch.autograd.grad(y, (x, m.weight), dy, retain_graph=True), warmup=3, iters=10)
        gf, gb = rows * hidden * 4 / 1e6 / f, rows * hidden * 6 / 1e6 / b
        print(json.dumps({"cfg": tag, "op": name, "h": hidden, "fwd_GBps": round(gf), "fwd_frac": round(gf / PEAK, 3), "bwd_GBps": round(gb),
                          "bwd_frac": round(gb / PEAK, 3)}), flush=True)
        del x, dy, y
