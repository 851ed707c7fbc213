#!/bin/bash
# GPU session A (1 GPU): GEMM / LN correctness + benches, then ncu --set full captures of the top kernels.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_layer_norm.py -q -x 2>&1 | tail -8
timeout 200 python benchmarks/bench_gemm.py 2>&1 | tail -9 | cut -c1-330
timeout 200 python benchmarks/bench_ops.py --what norm 2>&1 | grep -E "bwd|fwd" | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:gemm2_kernel -s 2 -c 1 -o gpurun_out/ncu_gemm2 python benchmarks/profile_targets.py gemm > gpurun_out/ncu_gemm2.log 2>&1
APEX_B200_GEMM_1CTA=1 timeout 300 $NCU -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/ncu_gemm1 python benchmarks/profile_targets.py gemm > gpurun_out/ncu_gemm1.log 2>&1
timeout 300 $NCU -k regex:dist_step_kernel -s 2 -c 1 -o gpurun_out/ncu_dist_adam python benchmarks/profile_targets.py dist_adam > gpurun_out/ncu_dist_adam.log 2>&1
timeout 300 $NCU -k regex:ln_ -s 4 -c 3 -o gpurun_out/ncu_layer_norm python benchmarks/profile_targets.py layer_norm > gpurun_out/ncu_layer_norm.log 2>&1
timeout 300 $NCU -k regex:mt_kernel -s 2 -c 1 -o gpurun_out/ncu_mt_adam python benchmarks/profile_targets.py adam > gpurun_out/ncu_mt_adam.log 2>&1
timeout 300 $NCU -k regex:syncbn -s 2 -c 4 -o gpurun_out/ncu_syncbn python benchmarks/profile_targets.py syncbn > gpurun_out/ncu_syncbn.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/ncu_*.log
