#!/bin/bash
mkdir -p gpurun_out/sanitizer
timeout 600 python -m pytest tests/test_gpu_group_norm.py tests/test_gpu_layer_norm.py -x -q 2>&1 | tail -2
timeout 900 compute-sanitizer --print-limit 20 --tool racecheck --racecheck-report all --kernel-name kns=ln_ python -m pytest tests/test_gpu_layer_norm.py -x -q -k "False-False-dtype2 and (4096 or 16384)" > gpurun_out/sanitizer/racecheck_ln_only.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer/racecheck_ln_only.log | tail -2
timeout 900 compute-sanitizer --print-limit 20 --tool racecheck --racecheck-report all --kernel-name kns=gn_group python -m pytest tests/test_gpu_group_norm.py -x -q -k "dtype1" > gpurun_out/sanitizer/racecheck_gn_cluster.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer/racecheck_gn_cluster.log | tail -2
grep -A2 "hazard\|invalid" gpurun_out/sanitizer/racecheck_gn_cluster.log | head -12
