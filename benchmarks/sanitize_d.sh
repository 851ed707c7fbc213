#!/bin/bash
# Second sanitizer sweep: racecheck on the remaining shared-memory kernels (restricted to this library's kernels by name), memcheck on the
# attention backward / optimizers / SyncBN single-GPU paths.
mkdir -p gpurun_out/sanitizer
rc() { name=$1; kns=$2; shift 2; timeout 900 compute-sanitizer --print-limit 10 --tool racecheck --racecheck-report all --kernel-name kns=$kns "$@" > gpurun_out/sanitizer/$name.log 2>&1; echo "== $name: $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitizer/$name.log | tail -1) $(grep -E ' passed| failed' gpurun_out/sanitizer/$name.log | tail -1)"; grep -A2 "hazard detected\|invalid" gpurun_out/sanitizer/$name.log | grep " at " | sed -E 's/\+0x[0-9a-f]+//; s/.* at //' | cut -c1-110 | sort | uniq -c | sort -rn | head -4; }
mc() { name=$1; shift; timeout 900 compute-sanitizer --print-limit 10 --tool memcheck "$@" > gpurun_out/sanitizer/$name.log 2>&1; echo "== $name: $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer/$name.log | tail -1) $(grep -E ' passed| failed' gpurun_out/sanitizer/$name.log | tail -1)"; }
rc racecheck_softmax_xent softmax python -m pytest tests/test_gpu_softmax_xent_rope.py -x -q -k "causal_softmax or scaled_masked_softmax"
rc racecheck_xent xentropy python -m pytest tests/test_gpu_softmax_xent_rope.py -x -q -k "test_xentropy"
rc racecheck_mt mt_kernel python -m pytest tests/test_gpu_multi_tensor.py -x -q -k "l2norm or lamb"
rc racecheck_syncbn syncbn python -m pytest tests/test_gpu_syncbn.py -x -q -k "single or world1 or one_gpu or channels_last"
rc racecheck_transducer kernel python -m pytest tests/test_gpu_contrib.py -x -q -k "transducer_loss or focal"
rc racecheck_conv_epi conv_epi python -m pytest tests/test_gpu_contrib.py -x -q -k "conv_epilogue"
mc memcheck_fmha_bwd python -m pytest tests/test_gpu_fmha.py -x -q -k "bwd_fixed_length and 200 or rescaling and ramp"
mc memcheck_optim python -m pytest tests/test_gpu_optimizers.py tests/test_gpu_multi_tensor.py -x -q -k "adam or sgd"
mc memcheck_syncbn python -m pytest tests/test_gpu_syncbn.py -x -q -k "not gpus"
