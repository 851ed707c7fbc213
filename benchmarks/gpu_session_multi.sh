#!/bin/bash
# Multi-GPU session: usage gpu_session_multi.sh N [tests]   (run under `gpurun --gpus N`)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests/test_gpu_dist_adam.py tests/test_gpu_syncbn.py tests/test_gpu_contrib.py -q -x -k "gpus or halo" 2>&1 | tail -8
fi
echo "== ours (policy auto)"; timeout 600 $TR --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n$N.json | cut -c1-900
echo "== ours P2P"; APEX_B200_DIST_NVLS=0 timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_ours_p2p_n$N.json | cut -c1-400
echo "== ours NVLS"; APEX_B200_DIST_NVLS=1 timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_ours_nvls_n$N.json | cut -c1-400
echo "== reference"; timeout 900 $TR --master-port 29614 bench.py --impl reference --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_n$N.json | cut -c1-600
echo "== syncbn resnet50"; timeout 600 $TR --master-port 29615 benchmarks/bench_syncbn.py --steps 10 --warmup 4 2>&1 | tail -4 | tee gpurun_out/bench_syncbn_n$N.txt | cut -c1-1500
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw --format=csv | head -9
