#!/bin/bash
# GPU session C (1 GPU): correctness of the retuned GEMM / LayerNorm kernels and of the new contrib kernels, then benches + ncu.
mkdir -p gpurun_out/ncu
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_layer_norm.py tests/test_gpu_contrib.py tests/test_gpu_syncbn.py tests/test_gpu_softmax_xent_rope.py -q -x 2>&1 | tail -15
timeout 200 python benchmarks/bench_gemm.py 2>&1 | tail -9 | cut -c1-330
timeout 200 python benchmarks/bench_ops.py --what norm 2>&1 | grep -E "bwd|fwd" | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4 tgt=$5
  timeout 300 env $6 $NCU -k regex:$rx -s $skip -c $cnt -o gpurun_out/ncu/$name python benchmarks/profile_targets.py $tgt > gpurun_out/ncu/$name.log 2>&1
  ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu/$name.source.csv.gz
  tail -1 gpurun_out/ncu/$name.log
}
cap gemm2 gemm2_kernel 2 1 gemm
cap ln_fwd ln_fwd_vec 1 1 layer_norm
cap ln_bwd ln_bwd_vec 1 1 layer_norm
rm -f gpurun_out/ncu/ln_fwd.ncu-rep gpurun_out/ncu/ln_bwd.ncu-rep
du -sh gpurun_out
