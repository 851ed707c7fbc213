#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
for MB in 0 100000; do
  APEX_B200_GN_STREAM_MIN_MB=$MB timeout 600 python benchmarks/bench_group_norm.py 2>&1 | grep "^{" > gpurun_out/bench_group_norm_thr$MB.json
done
python - <<'PY'
import json
def load(f):
    out={}
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        rows=d if isinstance(d,list) else d.get("rows",[d])
        for r in rows:
            if isinstance(r,dict) and "HW" in r: out[(r["G"],r["HW"],r["C"])]=r
    return out
a,b=load("gpurun_out/bench_group_norm_thr0.json"),load("gpurun_out/bench_group_norm_thr100000.json")
for k in sorted(a):
    s,o=a[k],b.get(k,{})
    print(k, "MB", round(8*k[1]*k[2]*2/1e6,1), "fwd stream/slab/ref", s.get("ours_fwd_us"), o.get("ours_fwd_us"), s.get("reference_fwd_us"), "| bwd", s.get("ours_bwd_us"), o.get("ours_bwd_us"), s.get("reference_bwd_us"))
PY
