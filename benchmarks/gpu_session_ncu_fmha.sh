#!/bin/bash
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
name=fmha_fwd_online
timeout 300 $NCU -k "regex:fmha_fwd_kernel" -s 1 -c 1 -o gpurun_out/ncu/$name python benchmarks/profile_targets.py fmha > gpurun_out/ncu/$name.log 2>&1
ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu/$name.source.csv.gz
rm -f gpurun_out/ncu/$name.ncu-rep
python benchmarks/ncu_source_top.py $name 25
grep -E "sm__pipe_tensor_cycles_active.avg.pct|smsp__issue_active.avg.pct|sm__throughput" gpurun_out/ncu/$name.raw.csv | head -3 | cut -c1-200
