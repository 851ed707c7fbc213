#!/bin/bash
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node $N"
echo "== SyncBN ResNet-50 N=$N"
timeout 400 $TR benchmarks/bench_syncbn.py --steps 12 --warmup 4 2>&1 | grep "^{" | tee gpurun_out/bench_syncbn_n$N.json | cut -c1-600
echo "== ours N=$N e2e with overlap_grad_sync off"
timeout 400 $TR bench.py --gpus $N --steps 6 --warmup 3 --overlap off 2>&1 | tail -1 > gpurun_out/bench_ours_n${N}_nooverlap.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ours_n${N}_nooverlap.json").read().strip().splitlines()[-1])
print("no-overlap: step", round(d["value"],2), "e2e", round(d["e2e"]["value"],2))
PY
echo "== tests (halo exchange device epoch, pool peer map, spatial bottleneck) at 2 ranks"
timeout 500 python -m pytest tests/test_gpu_contrib.py -m gpu -q -k "halo or peer_map or spatial or nccl_p2p" 2>&1 | tail -5 | cut -c1-250
