#!/bin/bash
export CUDA_LAUNCH_BLOCKING=1
for f in test_gpu_group_norm test_gpu_syncbn test_gpu_contrib; do
  echo "=== $f"; timeout 600 python -m pytest tests/$f.py -x -q 2>&1 | grep -v "^$" | tail -40 | cut -c1-220
done
