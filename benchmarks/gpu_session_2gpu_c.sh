#!/bin/bash
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node $N"
timeout 1200 python -m pytest tests/test_gpu_dist_adam.py tests/test_gpu_contrib.py tests/test_gpu_syncbn.py -m gpu -q -k "two_gpus or world1_step_reads or world1_dist_lamb or nccl_p2p or spatial_bottleneck or syncbn" 2>&1 | tail -120 | cut -c1-250 > gpurun_out/tests_2gpu_c.log; tail -12 gpurun_out/tests_2gpu_c.log
echo "== ours N=2"
timeout 500 $TR bench.py --gpus $N --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n$N.json | cut -c1-200
python - <<'PY'
import json
for f in ("gpurun_out/bench_ours_n2.json",):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"],2), round(d["sequence_ms_per_step"],2), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.get("e2e",{}).items() if k in("value","gpu_launches","step_in_backward")})
PY
