#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group_norm.py -x -q 2>&1 | tail -2
APEX_B200_GN_STREAM_MIN_MB=0 timeout 600 python -m pytest tests/test_gpu_group_norm.py -x -q 2>&1 | tail -2
for thr in default 0 30; do
  if [ $thr = default ]; then unset APEX_B200_GN_STREAM_MIN_MB; else export APEX_B200_GN_STREAM_MIN_MB=$thr; fi
  timeout 900 python benchmarks/bench_group_norm.py > gpurun_out/bench_group_norm_$thr.log 2>&1
  cp gpurun_out/bench_group_norm.json gpurun_out/bench_group_norm_thr_$thr.json
done
