#!/bin/bash
# Headline bench (both arms) at N = 2 on the final tree (the multi-GPU tests ran in the previous session: 20 passed).
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== reference N=2"; timeout 600 $TR --master-port 29631 bench.py --impl reference --gpus $N --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_n$N.json | cut -c1-160
echo "== ours N=2"; timeout 600 $TR --master-port 29632 bench.py --gpus $N --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_n$N.json | cut -c1-160
python - <<'PY'
import json
for f in ("gpurun_out/bench_ref_n2.json", "gpurun_out/bench_ours_n2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "step", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d.get("clocks", {}).get("reasons"))
    except Exception as e:
        print(f, "unreadable", e)
PY
