#!/bin/bash
mkdir -p gpurun_out/sanitizer
CS="compute-sanitizer --print-limit 3000"
for t in test_world1_fused_kernel_matches_adamw test_world1_capturable_and_state_dict test_world1_overlap_param_sync_into_fused_dense test_world1_step_reads_handed_over_gradients_in_place test_world1_step_in_backward_matches_step_after_backward test_world1_cuda_graph_capture test_world1_dist_lamb; do
  timeout 600 $CS --tool initcheck python -m pytest tests/test_gpu_dist_adam.py -x -q -k "$t" > gpurun_out/sanitizer/init_$t.log 2>&1
  echo "== $t: $(grep -c 'Uninitialized' gpurun_out/sanitizer/init_$t.log) reports; kernels:"
  grep -A1 "Uninitialized" gpurun_out/sanitizer/init_$t.log | grep " at " | sed -E 's/\+0x[0-9a-f]+//; s/.* at //' | cut -c1-150 | sort | uniq -c | sort -rn | head -5
done
timeout 600 compute-sanitizer --print-limit 50 --tool racecheck --racecheck-report all --kernel-regex kns=ln_ python -m pytest tests/test_gpu_layer_norm.py -x -q -k "False-False-dtype2 and (4096 or 16384)" > gpurun_out/sanitizer/racecheck_ln_only.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer/racecheck_ln_only.log | tail -3
grep -A3 "hazard detected" gpurun_out/sanitizer/racecheck_ln_only.log | grep " at " | sed -E 's/\+0x[0-9a-f]+//; s/.* at //' | cut -c1-120 | sort | uniq -c | sort -rn | head -6
