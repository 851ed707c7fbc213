#!/bin/bash
# compute-sanitizer runs (one GPU under gpurun; each tool slows kernels 10-100x, so the selections are small). Logs go to gpurun_out/sanitizer/.
# The reference has no sanitizer configuration at all; the cross-CTA / cross-GPU protocols here (grid barriers,
# arrival counters, cluster exchanges, epoch flags) are exactly the kind of code racecheck / synccheck are for.
mkdir -p gpurun_out/sanitizer
CS="compute-sanitizer --error-exitcode 1 --print-limit 5"
run() { name=$1; shift; echo "== $name: $*"; timeout 900 "$@" > gpurun_out/sanitizer/$name.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer/$name.log | tail -3; }
run memcheck_layer_norm $CS --tool memcheck python -m pytest tests/test_gpu_layer_norm.py -x -q -k "dtype2 and (4096 or 16384 or 12288 or 100)"
run memcheck_gemm_tf32 $CS --tool memcheck python -m pytest tests/test_gpu_gemm.py -x -q -k "tf32_layouts and (200 or 260) or tf32_epilogues"
run memcheck_group_norm $CS --tool memcheck python -m pytest tests/test_gpu_group_norm.py -x -q -k "dtype1"
run memcheck_conv_fmha $CS --tool memcheck python -m pytest tests/test_gpu_contrib.py tests/test_gpu_fmha.py -x -q -k "conv or (fixed_length and True)"
run racecheck_layer_norm $CS --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_layer_norm.py -x -q -k "False-False-dtype2 and (4096 or 16384)"
run synccheck_cluster_bwd $CS --tool synccheck python -m pytest tests/test_gpu_layer_norm.py tests/test_gpu_group_norm.py -x -q -k "dtype2-shape10 or dtype2-shape4 or (silu and dtype1)"
run initcheck_dist_adam $CS --tool initcheck python -m pytest tests/test_gpu_dist_adam.py -x -q -k "world1 or single"
