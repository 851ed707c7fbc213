#!/bin/bash
# 8-GPU confirmation (run under gpurun --gpus 8; every minute here costs 8): headline bench both arms, link probes, exposed-communication
# benchmark, the two 8-GPU tests.
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out profiles/results
TR="python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node $N"
echo "== ours N=$N"
timeout 400 $TR bench.py --gpus $N --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n$N.json | cut -c1-250
echo "== reference N=$N"
timeout 500 $TR bench.py --impl reference --gpus $N --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_n$N.json | cut -c1-250
echo "== link probes"
timeout 200 $TR benchmarks/bench_symm.py --mb 2048 --iters 5 --ctas 148 296 --unroll 2 8 --write-peaks 2>&1 | grep "^{" > gpurun_out/symm_n$N.jsonl; tail -2 gpurun_out/symm_n$N.jsonl | cut -c1-500
cp profiles/results/link_peaks.json gpurun_out/link_peaks_n$N.json 2>/dev/null
echo "== llama block: exposed communication"
timeout 300 $TR benchmarks/bench_llama_block.py --layers 8 --tokens 4096 2>&1 | grep "^{" | tee gpurun_out/llama_block_n$N.json | cut -c1-700
echo "== tests at $N GPUs"
timeout 400 python -m pytest tests/test_gpu_dist_adam.py tests/test_gpu_syncbn.py -m gpu -q -x -k "eight or 8" 2>&1 | tail -4 | cut -c1-250
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*_n8.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("impl"), round(d["value"],2), round(d["sequence_ms_per_step"],2), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.get("e2e",{}).items() if k in("value","unavailable")})
    except Exception as e:
        print(f, "unreadable", e)
PY
