#!/bin/bash
# GPU session E (1 GPU): validate the dependency-driven GroupNorm, SyncBN merge fix, LN fwd; head-to-head benches vs the reference.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layer_norm.py tests/test_gpu_syncbn.py tests/test_gpu_group_norm.py tests/test_gpu_contrib.py -q 2>&1 | tail -12
echo "== LN"; timeout 200 python benchmarks/bench_ops.py --what norm 2>&1 | grep -E "LayerNorm (bwd|fwd)|RMSNorm (bwd|fwd)" | grep -v torch | cut -c1-140
echo "== group norm"; GN_BATCH=8 timeout 300 python benchmarks/bench_group_norm.py 2>&1 | cut -c1-330
echo "== syncbn N=1"; timeout 300 python benchmarks/bench_syncbn.py --steps 10 --warmup 4 2>&1 | tail -2 | cut -c1-1500
echo "== vs reference"; timeout 600 python benchmarks/bench_vs_reference.py 2>&1 | cut -c1-300
echo "== permutation"; timeout 300 python benchmarks/bench_permutation.py 2>&1 | cut -c1-300
