#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_contrib.py tests/test_gpu_dist_adam.py tests/test_gpu_gemm.py -m gpu -q -x -k "conv_epilogue or bottleneck or world1 or bias_gradient" 2>&1 | tail -30 | cut -c1-300 | tee gpurun_out/tests_1gpu_b.log
echo "== ours N=1 default"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1.json | cut -c1-300
echo "== ours N=1 step-in-backward"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --step-in-backward on 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_sib.json | cut -c1-300
python - <<'PY'
import json
for f in ("gpurun_out/bench_ours_n1.json","gpurun_out/bench_ours_n1_sib.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"],2), round(d["sequence_ms_per_step"],2), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.get("e2e",{}).items() if k in("value","gpu_launches","step_in_backward")})
PY
