#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29615 benchmarks/bench_syncbn.py --steps 8 --warmup 3 > gpurun_out/syncbn_n$N.log 2>&1
echo "exit $?"
grep -n -i -E "error|apex_b200|trap|abort|assert|Traceback|never|timeout" gpurun_out/syncbn_n$N.log | head -30 | cut -c1-300
tail -5 gpurun_out/syncbn_n$N.log | cut -c1-1500
