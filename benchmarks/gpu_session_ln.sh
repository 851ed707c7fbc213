#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/ln_sweep.jsonl
timeout 600 python -m pytest tests/test_gpu_layer_norm.py -x -q 2>&1 | tail -3
run() { tag=$1; shift; env "$@" timeout 200 python benchmarks/bench_ln_sweep.py $tag ${H:-1024,2048,4096,8192,12288,16384} 2>&1 | grep '^{' | tee -a gpurun_out/ln_sweep.jsonl; }
H=12288,16384 run cluster_push A=1
H=12288,16384 run nocluster APEX_B200_LN_BWD_CLUSTER=0
