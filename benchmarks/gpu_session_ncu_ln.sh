#!/bin/bash
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
cap() { name=$1; rx=$2; shift 2
  env "$@" timeout 300 $NCU -k "regex:$rx" -s 1 -c 1 -o gpurun_out/ncu/$name python benchmarks/profile_targets.py layer_norm > gpurun_out/ncu/$name.log 2>&1
  ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu/$name.source.csv.gz
  rm -f gpurun_out/ncu/$name.ncu-rep
  echo "=== $name"; python benchmarks/ncu_source_top.py $name 22
  python - $name <<'PY'
import csv, sys
rows = list(csv.reader(open(f"gpurun_out/ncu/{sys.argv[1]}.raw.csv")))
hdr = rows[0]; vals = rows[-1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_uniform.sum"]
for w in want:
    if w in hdr: print(f"{w:95s} {vals[hdr.index(w)]}")
PY
}
cap ln_fwd_col4096 "ln_fwd_col" A=1
cap ln_fwd_old4096 "ln_fwd_vec" APEX_B200_LN_FWD_COL=0
