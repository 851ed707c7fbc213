"""Top stall-sampled SASS instructions of an `ncu --page source --csv` export. usage: ncu_source_top.py <name> [N]"""
import csv, gzip, io, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(io.TextIOWrapper(gzip.open(os.path.join(ROOT, 'gpurun_out/ncu', name + '.source.csv.gz')))))
# a file may contain several kernels: sections start with a header row
secs = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r]
for si, hi in enumerate(secs):
    hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
    end = secs[si + 1] if si + 1 < len(secs) else len(rows)
    body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
    sc = col['# Samples']
    tot = sum(int(r[sc] or 0) for r in body)
    print(f"--- kernel section {si}: {len(body)} instructions, {tot} samples")
    stalls = [h for h in hdr if h.startswith('stall_')]
    for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][sc] or 0))[:N]:
        n = int(r[sc] or 0)
        if n == 0: break
        st = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"{idx:5d} {100.0*n/max(tot,1):5.1f}%  {r[col['Source']][:90]:90s} {st}")
