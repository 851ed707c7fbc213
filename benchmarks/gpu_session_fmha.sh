#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fmha.py tests/test_gpu_contrib.py -m gpu -q -x -k "fmha or multihead or attn" 2>&1 | tail -4 | cut -c1-300
timeout 600 python benchmarks/bench_fmha.py 2>&1 | grep "^{" > gpurun_out/bench_fmha.json; python - <<'PY'
import json
for l in open("gpurun_out/bench_fmha.json"):
    d=json.loads(l)
    if "summary" in d: print(d)
    else: print(d["d"], d["causal"], d["seq"], "fwd", round(d["ours_fwd_ms"],3), round(d["sdpa_fwd_ms"],3), "f+b", round(d["ours_fwd_bwd_ms"],3), round(d["sdpa_fwd_bwd_ms"],3), "TF", round(d["ours_fwd_tflops"]))
PY
