#!/bin/bash
# 2-GPU validation: multi-GPU tests (dist adam, syncbn, halo exchange, lamb) + GN after the cluster change
timeout 300 python -m pytest tests/test_gpu_group_norm.py -x -q 2>&1 | tail -3 | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_dist_adam.py tests/test_gpu_syncbn.py tests/test_gpu_contrib.py -q -x -k "gpus or halo" 2>&1 | tail -12 | cut -c1-250
echo "== group norm"; GN_BATCH=8 timeout 300 python benchmarks/bench_group_norm.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(r['G'], r['act'], r['HW'], r['C'], 'fwd', r['ours_fwd_us'], r.get('reference_fwd_us'), 'bwd', r['ours_bwd_us'], r.get('reference_bwd_us'))
"
