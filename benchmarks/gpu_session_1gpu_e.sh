#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fmha.py tests/test_gpu_softmax_xent_rope.py tests/test_gpu_contrib.py tests/test_gpu_ext_compat.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-300
echo "== fmha vs sdpa"
timeout 600 python benchmarks/bench_fmha.py 2>&1 | grep "^{" > gpurun_out/bench_fmha.json; python - <<'PY'
import json
for l in open("gpurun_out/bench_fmha.json"):
    d=json.loads(l)
    if "summary" in d: print(d)
    else: print(d["d"], d["causal"], d["seq"], "fwd", round(d["ours_fwd_ms"],3), round(d["sdpa_fwd_ms"],3), "f+b", round(d["ours_fwd_bwd_ms"],3), round(d["sdpa_fwd_bwd_ms"],3), "TF", round(d["ours_fwd_tflops"]))
PY
echo "== W1 sweep"; timeout 600 python benchmarks/bench_w1_variants.py 2>&1 | tee gpurun_out/w1_variants_b.jsonl
echo "== MT grid"; timeout 900 python benchmarks/bench_mt_grid.py 2>&1 | tee gpurun_out/mt_grid.jsonl
echo "== rope"; timeout 200 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from apex_b200.transformer.functional import fused_apply_rotary_pos_emb
from apex_b200.utils.timing import time_fn
t = torch.randn(4096, 8, 32, 128, device="cuda", dtype=torch.bfloat16)
freqs = torch.randn(4096, 1, 1, 128, device="cuda")
with torch.no_grad():
    med, mn = time_fn(lambda: fused_apply_rotary_pos_emb(t, freqs), 3, 10)
print("rope fwd ms", med, "GB/s", t.numel() * 4 / med / 1e6)
PY
