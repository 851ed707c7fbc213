#!/bin/bash
timeout 200 python benchmarks/bench_gemm_fp8.py 2>&1 | tail -1 | cut -c1-500
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_layer_norm.py -x -q 2>&1 | tail -2 | cut -c1-200
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 200 python benchmarks/bench_ops.py --what norm 2>&1 | grep -E "Norm fwd" | grep -v torch | cut -c1-140
