"""Registers, spills (STACK bytes) and static shared memory of EVERY kernel in the built objects, from `cuobjdump --dump-resource-usage`
(what `-Xptxas -v` prints at compile time), per translation unit. Writes profiles/resource_usage.md. Needs no GPU.
usage: python benchmarks/resource_usage.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build", "obj")


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, timeout=120).stdout.splitlines()
        return out if len(out) == len(names) else names
    except Exception:  # noqa: BLE001
        return names


def short(name, width=150):
    name = name.replace("void ", "").replace("ab::", "").replace("(anonymous namespace)::", "")
    cut = name.rfind(">(")                       # keep the template arguments, drop the parameter list
    name = name[:cut + 1] if cut >= 0 else name.split("(")[0]
    return name if len(name) <= width else name[:width - 3] + "..."


def ptxas_spills(units):
    """{unit: {mangled kernel: (stack, spill stores, spill loads)}} from `nvcc -Xptxas -v` with the build's own flags (recompiles the listed
    translation units into a scratch directory: minutes)."""
    import concurrent.futures
    import importlib.util
    import tempfile

    spec = importlib.util.spec_from_file_location("apex_b200_build", os.path.join(ROOT, "apex_b200", "_build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    tmp = tempfile.mkdtemp(prefix="ptxas_")

    def one(unit):
        src = os.path.join(ROOT, "apex_b200", "csrc", unit + ".cu")
        flags = b.NVCC_FLAGS + ["-I", str(b.CSRC)] + (b._cutlass_include() if "cute" in open(src).read() or "cutlass" in open(src).read() else [])
        log = subprocess.run([b.NVCC] + flags + ["-Xptxas", "-v", "-c", src, "-o", os.path.join(tmp, unit + ".o")],
                             capture_output=True, text=True, timeout=3600).stderr
        res = {}
        for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log):
            res[m.group(1)] = tuple(int(m.group(i)) for i in (2, 3, 4))
        return unit, res

    with concurrent.futures.ThreadPoolExecutor(4) as ex:
        return dict(ex.map(one, units))


def main():
    import sys

    rows, spills = [], []
    want_ptxas = "--ptxas" in sys.argv
    for f in sorted(os.listdir(OBJ)):
        if not f.endswith(".o"):
            continue
        txt = subprocess.run(["cuobjdump", "--dump-resource-usage", os.path.join(OBJ, f)], capture_output=True, text=True, timeout=300).stdout
        fn = re.findall(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", txt)
        if not fn:
            continue
        names = demangle([x[0] for x in fn])
        regs = [int(x[1]) for x in fn]
        hist = collections.Counter("<=32" if r <= 32 else "<=64" if r <= 64 else "<=128" if r <= 128 else "<=168" if r <= 168 else ">168" for r in regs)
        top = max(range(len(fn)), key=lambda i: regs[i])
        spilled = [(names[i], int(fn[i][2]), regs[i], fn[i][0]) for i in range(len(fn)) if int(fn[i][2]) > 0]
        spills += [(f, *s) for s in spilled]
        rows.append(f"| {f[:-2]}.cu | {len(fn)} | {hist['<=32']} / {hist['<=64']} / {hist['<=128']} / {hist['<=168']} / {hist['>168']} | {regs[top]} (`{short(names[top], 70)}`) | "
                    f"{max(int(x[3]) for x in fn)} | {len(spilled)} |")
    out = ["# Per-kernel resource usage (cuobjdump --dump-resource-usage of build/obj/*.o, sm_100a; generator `benchmarks/resource_usage.py`)", "",
           "Registers per thread decide how many CTAs fit on an SM (64 K registers: 1024-thread residency needs <= 64, 512 threads <= 128); STACK > 0 "
           "means the compiler spilled registers to local memory. Static shared memory only (dynamic shared memory is set at launch).", "",
           "| translation unit | kernels | registers <=32 / <=64 / <=128 / <=168 / >168 | most registers (kernel) | max static smem (B) | kernels with a stack frame |",
           "|---|---|---|---|---|---|"] + rows
    total = sum(int(r.split('|')[2]) for r in rows)
    out += ["", f"## Kernels with a stack frame ({len(spills)} of {total})", "",
            "A stack frame holds local arrays the kernel indexes dynamically (peer-pointer tables, per-thread staging) and, when ptxas runs out of "
            "registers under a `__launch_bounds__` cap, spilled registers. The two are told apart by `nvcc -Xptxas -v` (`--ptxas`):", ""]
    if want_ptxas:
        info = ptxas_spills(sorted({f[:-2] for f, *_ in spills}))
        out += ["| translation unit | kernels with a stack frame | of which spill registers | largest spill (kernel: stack / spill stores / spill loads, registers) |",
                "|---|---|---|---|"]
        by_unit = collections.defaultdict(list)
        for f, name, stack, regs, mangled in spills:
            by_unit[f[:-2]].append((name, regs, info.get(f[:-2], {}).get(mangled, (stack, -1, -1))))
        for unit, items in sorted(by_unit.items()):
            sp = [x for x in items if x[2][1] > 0]
            worst = max(sp, key=lambda x: x[2][1], default=None)
            desc = f"`{short(worst[0], 90)}`: {worst[2][0]} / {worst[2][1]} / {worst[2][2]} B, {worst[1]} registers" if worst else "none: local arrays only"
            out.append(f"| {unit}.cu | {len(items)} | {len(sp)} | {desc} |")
        out += ["", "### Every kernel that spills", "", "| translation unit | kernel | stack B | spill stores B | spill loads B | registers |", "|---|---|---|---|---|---|"]
        for unit, items in sorted(by_unit.items()):
            for name, regs, (stack, st, ld) in sorted(items, key=lambda x: -x[2][1]):
                if st > 0:
                    out.append(f"| {unit}.cu | `{short(name)}` | {stack} | {st} | {ld} | {regs} |")
    else:
        out += ["| translation unit | kernel | stack bytes | registers |", "|---|---|---|---|"]
        out += [f"| {f[:-2]}.cu | `{short(n)}` | {s} | {r} |" for f, n, s, r, _ in sorted(spills, key=lambda x: -x[2])]
    open(os.path.join(ROOT, "profiles", "resource_usage.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:40]))
    print(f"... {len(spills)} kernels with a stack frame")


if __name__ == "__main__":
    main()
