#!/bin/bash
# GPU session B (1 GPU): ncu --set full captures of the top kernels; reports are exported to CSV on the box and only the two most
# important .ncu-rep files are kept (gpurun_out/ must stay under 64 MiB).
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, skip, count, target, [env]
  local name=$1 rx=$2 skip=$3 cnt=$4 tgt=$5
  timeout 300 env $6 $NCU -k regex:$rx -s $skip -c $cnt -o gpurun_out/ncu/$name python benchmarks/profile_targets.py $tgt > gpurun_out/ncu/$name.log 2>&1
  ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$name.ncu-rep --page details --csv > gpurun_out/ncu/$name.details.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$name.ncu-rep --page source --csv > gpurun_out/ncu/$name.source.csv 2>/dev/null
  tail -2 gpurun_out/ncu/$name.log
}
cap gemm2 gemm2_kernel 2 1 gemm
cap gemm1 gemm_kernel 2 1 gemm APEX_B200_GEMM_1CTA=1
cap dist_adam dist_step_kernel 2 1 dist_adam
cap ln_fwd ln_fwd_vec 1 1 layer_norm
cap ln_bwd ln_bwd_vec 1 1 layer_norm
cap mt_adam mt_kernel 2 1 adam
cap syncbn syncbn_kernel 4 4 syncbn
cap group_norm group_norm_kernel 2 2 group_norm
cap softmax softmax 1 1 softmax
cap xent xentropy_ 2 2 xent
cd gpurun_out/ncu
rm -f gemm1.ncu-rep ln_fwd.ncu-rep ln_bwd.ncu-rep mt_adam.ncu-rep syncbn.ncu-rep group_norm.ncu-rep softmax.ncu-rep xent.ncu-rep
gzip -f *.source.csv
du -sh . ; ls -la
