#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -k "fp8" 2>&1 | tail -12 | cut -c1-250
timeout 200 python benchmarks/bench_gemm_fp8.py 2>&1 | tail -4 | cut -c1-500
