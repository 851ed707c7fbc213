#!/bin/bash
# GPU session D (1 GPU): correctness after the syncbn / group_norm / LN / GEMM-epilogue rewrites, then benches.
mkdir -p gpurun_out/ncu
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_layer_norm.py tests/test_gpu_syncbn.py tests/test_gpu_group_norm.py tests/test_gpu_contrib.py -q 2>&1 | tail -25
timeout 200 python benchmarks/bench_gemm.py 2>&1 | tail -9 | cut -c1-330
echo "== LN target_v=2"; timeout 200 python benchmarks/bench_ops.py --what norm 2>&1 | grep -E "LayerNorm (bwd|fwd)|RMSNorm (bwd|fwd)" | grep -v torch | cut -c1-140
echo "== LN fwd target_v=4"; APEX_B200_LN_FWD_V=4 timeout 200 python benchmarks/bench_ops.py --what norm 2>&1 | grep -E "Norm fwd" | grep -v torch | cut -c1-140
echo "== group norm"; GN_BATCH=8 timeout 300 python benchmarks/bench_group_norm.py 2>&1 | cut -c1-400
echo "== syncbn N=1"; timeout 300 python benchmarks/bench_syncbn.py --steps 10 --warmup 4 2>&1 | tail -3 | cut -c1-1200
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4 tgt=$5
  timeout 300 env $6 $NCU -k regex:$rx -s $skip -c $cnt -o gpurun_out/ncu/$name python benchmarks/profile_targets.py $tgt > gpurun_out/ncu/$name.log 2>&1
  ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu/$name.source.csv.gz
  rm -f gpurun_out/ncu/$name.ncu-rep
  tail -1 gpurun_out/ncu/$name.log
}
cap syncbn syncbn_kernel 4 4 syncbn
cap group_norm group_norm_kernel 2 2 group_norm
cap ln_fwd ln_fwd_vec 1 1 layer_norm
du -sh gpurun_out
