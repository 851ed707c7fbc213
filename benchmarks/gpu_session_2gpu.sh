#!/bin/bash
# 2-GPU session: the whole GPU suite (multi-GPU cases run at 2 ranks), headline bench at N=2 (both arms, + step-in-backward variant).
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node $N"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | cut -c1-400
echo "== ours N=2"
timeout 500 $TR bench.py --gpus $N --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n$N.json | cut -c1-2200
echo "== ours N=2 step-in-backward"
timeout 500 $TR bench.py --gpus $N --steps 8 --warmup 3 --step-in-backward on 2>&1 | tail -1 | tee gpurun_out/bench_ours_n${N}_sib.json | cut -c1-2200
echo "== reference N=2"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_n$N.json | cut -c1-2200
