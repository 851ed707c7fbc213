#!/bin/bash
mkdir -p gpurun_out/ncu
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | cut -c1-300
echo "== W1 variants"
timeout 900 python benchmarks/bench_w1_variants.py 2>&1 | tee gpurun_out/w1_variants.jsonl
echo "== ncu (names with template arguments)"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
cap() {
  name=$1; target=$2; rx=$3; skip=${4:-1}
  timeout 240 $NCU -k "regex:$rx" -s $skip -c 1 -o gpurun_out/ncu/$name python benchmarks/profile_targets.py $target > gpurun_out/ncu/$name.log 2>&1
  if [ -f gpurun_out/ncu/$name.ncu-rep ]; then
    ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
    ncu -i gpurun_out/ncu/$name.ncu-rep --page source --csv 2>/dev/null | head -400 > gpurun_out/ncu/$name.source.csv
    rm -f gpurun_out/ncu/$name.ncu-rep; echo "captured $name"
  else echo "NO CAPTURE $name"; tail -2 gpurun_out/ncu/$name.log | cut -c1-200; fi
}
cap mt_lamb_stage1 lamb "LambStage1"
cap mt_lamb_stage2 lamb "LambStage2"
cap mt_sgd sgd "SgdOp"
cap mt_novograd novograd "NovoGrad"
cap mt_l2norm mt_basic "L2Norm"
cap mt_scale mt_basic "ScaleOp"
cap mt_axpby mt_basic "Axpby"
cap fmha_bwd_dkv fmha "fmha_bwd_kernel.*true"
cap fmha_bwd_dq fmha "fmha_bwd_kernel.*false"
cap conv_epilogue_bwd conv_epilogue "conv_epi_bwd"
cap conv_epilogue_fwd conv_epilogue "conv_epi_fwd"
