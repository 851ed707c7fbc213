#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_group_norm.py tests/test_gpu_syncbn.py -x -q 2>&1 | tail -6 | cut -c1-220
echo "== group norm"; GN_BATCH=8 timeout 300 python benchmarks/bench_group_norm.py 2>&1 | cut -c1-330
