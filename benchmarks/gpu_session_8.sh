#!/bin/bash
# 8-GPU session: one correctness test of the fused ZeRO step at 8 ranks, the flagship bench (ours, reference), SyncBN ResNet-50.
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_dist_adam.py -q -x -k "eight" 2>&1 | tail -3 | cut -c1-250
echo "== ours (policy auto)"; timeout 400 $TR --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n$N.json | cut -c1-1100
echo "== reference"; timeout 600 $TR --master-port 29614 bench.py --impl reference --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_n$N.json | cut -c1-700
echo "== syncbn resnet50"; timeout 400 $TR --master-port 29615 benchmarks/bench_syncbn.py --steps 8 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_syncbn_n$N.txt | cut -c1-1600
echo "== ours P2P"; APEX_B200_DIST_NVLS=0 timeout 300 $TR --master-port 29612 bench.py --gpus $N --steps 6 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_ours_p2p_n$N.json | cut -c1-400
