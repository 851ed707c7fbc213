#!/bin/bash
mkdir -p gpurun_out
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size"
for cfg in "ours_slab A=1" "ours_stream APEX_B200_GN_STREAM_MIN_MB=0" "ref A=1"; do
  set -- $cfg; tag=$1; shift
  which=ours; [ $tag = ref ] && which=ref
  env "$@" timeout 300 ncu $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/gn_probe_$tag.csv python benchmarks/gn_kernels_probe.py $which > /dev/null 2>&1
  python - $tag <<'PY'
import csv, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/gn_probe_{tag}.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
by = {}
for r in rows[1:]:
    by.setdefault(r[ii], {"k": r[ki][:90]})[r[mi]] = r[vi]
ids = sorted(by, key=int)
print("==", tag, len(ids), "kernels; last iteration:")
for i in ids[-(len(ids) // 3):]:
    d = by[i]
    print(f"  {d['k']:90s} {d.get('gpu__time_duration.sum','?'):>10s} ns grid {d.get('launch__grid_size','?')} rd {d.get('dram__bytes_read.sum','?')} wr {d.get('dram__bytes_write.sum','?')}")
PY
done
