"""Executed-instruction mix by SASS opcode from an `ncu --page source --csv` export. usage: ncu_opmix.py <name> [section]"""
import csv, gzip, io, sys, os, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1]; want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(io.TextIOWrapper(gzip.open(os.path.join(ROOT, 'gpurun_out/ncu', name + '.source.csv.gz')))))
secs = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r]
hi = secs[want]; hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
end = secs[want + 1] if want + 1 < len(secs) else len(rows)
mix = collections.Counter()
for r in rows[hi + 1:end]:
    if len(r) != len(hdr): continue
    src = r[col['Source']].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    op = op.split('.')[0]
    mix[op] += int(r[col['Instructions Executed']] or 0)
tot = sum(mix.values())
print(f"total warp-instructions executed: {tot:,}")
for op, n in mix.most_common(22):
    print(f"  {op:12s} {n:14,d}  {100.0*n/tot:5.1f}%")
