#!/bin/bash
# ncu --set full of the kernels changed late in round 2: LayerNorm fwd / bwd (hand decode, raw-bit gamma staging), tf32 GEMM (K-major fwd and
# the MN-major / BASE32B wgrad), plus the new ext_compat GPU test.
mkdir -p gpurun_out/ncu
timeout 600 python -m pytest tests/test_gpu_ext_compat.py -x -q 2>&1 | tail -3
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
cap() {  # name target kernel-regex skip
  name=$1; target=$2; rx=$3; skip=${4:-1}
  timeout 300 $NCU -k "regex:$rx" -s $skip -c 1 -o gpurun_out/ncu/$name python benchmarks/profile_targets.py $target > gpurun_out/ncu/$name.log 2>&1
  if [ -f gpurun_out/ncu/$name.ncu-rep ]; then
    ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
    rm -f gpurun_out/ncu/$name.ncu-rep
    echo "captured $name"
  else
    echo "NO CAPTURE $name"; tail -3 gpurun_out/ncu/$name.log
  fi
}
cap ln_fwd_r2b layer_norm "ln_fwd_vec" 1
cap ln_bwd_r2b layer_norm "ln_bwd_vec" 1
cap gemm_tf32_fwd gemm_tf32 "gemm2_kernel<float" 2
cap gemm_tf32_wgrad gemm_tf32 "gemm2_kernel<float" 3
