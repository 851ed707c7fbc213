#!/bin/bash
# Single-GPU verification: the whole GPU test-suite, smoke(), the headline bench (both arms).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | cut -c1-300
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_n1.json | cut -c1-1500
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1.json | cut -c1-1800
