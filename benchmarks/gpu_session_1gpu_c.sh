#!/bin/bash
mkdir -p gpurun_out
timeout 600 python benchmarks/bench_fmha.py 2>&1 | grep "^{" > gpurun_out/bench_fmha.json; tail -2 gpurun_out/bench_fmha.json | cut -c1-400
bash benchmarks/gpu_session_ncu.sh 2>&1 | tail -40
