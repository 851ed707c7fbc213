#!/bin/bash
# compute-sanitizer recipes (run under gpurun on one GPU; each tool slows kernels 10-100x, so the test selection is small).
# The reference has no sanitizer configuration at all; the cross-CTA / cross-GPU protocols here
# (grid barriers, arrival counters, epoch flags) are exactly the kind of code racecheck / synccheck are for.
set -x
T="tests/test_gpu_multi_tensor.py tests/test_gpu_layer_norm.py -k 'bf16 or float32' -x -q"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_gemm.py -x -q -k "not mlp" 2>&1 | tail -5
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_group_norm.py tests/test_gpu_syncbn.py -x -q -k "not gpus" 2>&1 | tail -8
timeout 1200 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_layer_norm.py -x -q -k "4096" 2>&1 | tail -5
timeout 1200 compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_dist_adam.py -x -q -k "world1" 2>&1 | tail -5
