#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group_norm.py -x -q 2>&1 | tail -3
timeout 900 python benchmarks/bench_group_norm.py > gpurun_out/bench_group_norm.log 2>&1; tail -2 gpurun_out/bench_group_norm.log | cut -c1-300
