#!/bin/bash
# ncu --set full captures (one GPU, one kernel instance per capture) of the named hot kernels; exports CSV on the box, keeps no .ncu-rep.
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
cap() {  # name target kernel-regex skip
  name=$1; target=$2; rx=$3; skip=${4:-2}
  timeout 240 $NCU -k "regex:$rx" -s $skip -c 1 -o gpurun_out/ncu/$name python benchmarks/profile_targets.py $target > gpurun_out/ncu/$name.log 2>&1
  if [ -f gpurun_out/ncu/$name.ncu-rep ]; then
    ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
    ncu -i gpurun_out/ncu/$name.ncu-rep --page source --csv 2>/dev/null | head -400 > gpurun_out/ncu/$name.source.csv
    rm -f gpurun_out/ncu/$name.ncu-rep
    echo "captured $name"
  else
    echo "NO CAPTURE $name"; tail -3 gpurun_out/ncu/$name.log
  fi
}
cap gemm2_384 gemm "gemm2_kernel" 2
cap gemm_dgelu_bgrad dgelu_bgrad "gemm2_kernel" 1
cap gemm_wgrad_accum wgrad "gemm2_kernel" 1
cap fmha_fwd fmha "fmha_fwd_kernel" 1
cap fmha_bwd_dkv fmha "fmha_bwd_kernel.*true" 1
cap fmha_bwd_dq fmha "fmha_bwd_kernel.*false" 1
cap zero_step_world1 dist_adam "dist_step_kernel" 2
cap mt_lamb_stage1 lamb "LambStage1" 1
cap mt_lamb_stage2 lamb "LambStage2" 1
cap mt_sgd sgd "SgdOp" 1
cap mt_novograd novograd "NovoGrad" 1
cap mt_l2norm mt_basic "L2Norm" 1
cap mt_scale mt_basic "ScaleOp" 1
cap mt_axpby mt_basic "Axpby" 1
cap update_scale_hysteresis mt_basic "update_scale_hysteresis" 1
cap softmax_bwd softmax_bwd "softmax_bwd" 1
cap xent_bwd xent "xentropy_bwd" 1
cap rope_fwd rope "rope_kernel" 1
cap conv_epilogue_bwd conv_epilogue "conv_epi_bwd" 1
cap conv_epilogue_fwd conv_epilogue "conv_epi_fwd" 1
cap ln_fwd_now layer_norm "ln_fwd" 1
cap ln_bwd_now layer_norm "ln_bwd" 1
ls gpurun_out/ncu | head -80
