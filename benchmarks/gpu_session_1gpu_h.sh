#!/bin/bash
# Single-GPU verification after the decode / LayerNorm / GEMM changes: whole GPU suite, smoke, ops bench, headline bench (both arms).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | cut -c1-300
timeout 600 python benchmarks/bench_ops.py --what adam,lamb,norm --out gpurun_out/bench_ops.json > gpurun_out/bench_ops.log 2>&1; tail -3 gpurun_out/bench_ops.log | cut -c1-300
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_n1.json | cut -c1-400
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1.json | cut -c1-400
