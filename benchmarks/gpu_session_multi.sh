#!/bin/bash
# Multi-GPU session (run under gpurun --gpus N): distributed-optimizer tests, link probes, headline bench variants.
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out profiles/results
TR="python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node $N"
echo "== tests (overlap_grad_sync, CUDA-graph capture of the distributed step, NVLS all-reduce, $N GPUs)"
timeout 600 python -m pytest tests/test_gpu_dist_adam.py tests/test_gpu_fmha.py -m gpu -q -x -k "two_gpus_overlap or cuda_graph_capture_of_the_distributed or nvls_allreduce or four_gpus" 2>&1 | tail -6 | cut -c1-300
echo "== link probes"
timeout 300 $TR benchmarks/bench_symm.py --mb 1024 --iters 5 --write-peaks 2>&1 | grep "^{" | tee gpurun_out/symm_n$N.jsonl | tail -3 | cut -c1-600
cp profiles/results/link_peaks.json gpurun_out/link_peaks_n$N.json 2>/dev/null
for H in ${HYBRIDS:-1.0 0.75 0.6}; do
  echo "== ours N=$N hybrid=$H"
  APEX_B200_DIST_HYBRID=$H timeout 400 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_ours_n${N}_h$H.json | cut -c1-400
done
echo "== ours N=$N P2P only"
APEX_B200_DIST_NVLS=0 timeout 400 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_ours_n${N}_p2p.json | cut -c1-400
echo "== ours N=$N full (e2e)"
APEX_B200_DIST_HYBRID=${BEST_HYBRID:-1.0} timeout 500 $TR bench.py --gpus $N --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n$N.json | cut -c1-2500
echo "== reference N=$N"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_n$N.json | cut -c1-2500
