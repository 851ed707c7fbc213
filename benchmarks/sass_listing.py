"""Trimmed SASS listings of the hot loops of every named kernel (profiles/sass/*.sass), from the built objects — no GPU needed.

For each selected function of build/obj/*.o: find the loops (backward branches), keep the innermost loops that contain the instructions
that matter for that kernel (tcgen05.mma = UTCHMMA / UTCQMMA, TMA = UTMALDG, tcgen05.ld = LDTM, multimem = LDGMC / STG...SYS, wide
global loads / stores, MUFU for the softmax exponentials) and write them with their addresses, plus the register / stack usage.
usage: python benchmarks/sass_listing.py"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build", "obj")
OUT = os.path.join(ROOT, "profiles", "sass")

# object -> list of (regex on the demangled function name, short tag, marker mnemonics ranked)
SELECT = {
    "gemm_sm100.o": [(r"gemm2_kernel<__nv_bfloat16, 6, 0>", "gemm2_bf16", ["UTCHMMA", "UTMALDG", "LDTM"]),
                     (r"gemm2_kernel<__nv_bfloat16, 6, 1>", "gemm2_fp8", ["UTCQMMA", "UTMALDG", "LDTM"]),
                     (r"gemm2_kernel<float, 6, 2>", "gemm2_tf32", ["UTCHMMA", "UTMALDG", "LDTM"]),
                     (r"gemm_kernel<__nv_bfloat16, 256, 4>", "gemm1_bf16", ["UTCHMMA", "UTMALDG", "LDTM"])],
    "fmha_fwd_sm100.o": [(r"fmha_fwd_kernel<__nv_bfloat16, 128>", "fmha_fwd_d128", ["UTCHMMA", "UTMALDG", "MUFU.EX2", "LDTM"])],
    "fmha_bwd_sm100.o": [(r"fmha_bwd_kernel<__nv_bfloat16, 128, true>", "fmha_bwd_dkv_d128", ["UTCHMMA", "UTMALDG", "MUFU.EX2", "LDTM"]),
                         (r"fmha_bwd_kernel<__nv_bfloat16, 128, false>", "fmha_bwd_dq_d128", ["UTCHMMA", "MUFU.EX2", "LDTM"])],
    "dist_adam.o": [(r"dist_step_kernel<__nv_bfloat16, __nv_bfloat16, 0, true, 1, 8>", "zero_step_nvls", ["LDGMC", "STG", "LDG"]),
                    (r"dist_step_kernel<__nv_bfloat16, __nv_bfloat16, 0, false, 2, 4>", "zero_step_p2p2", ["LDG.E.NA", "STG.E.NA", "LDG"]),
                    (r"dist_step_kernel<__nv_bfloat16, __nv_bfloat16, 0, false, 1, 2>", "zero_step_world1", ["LDG", "STG"]),
                    (r"dist_step_kernel<__nv_bfloat16, __nv_bfloat16, 3, true, 1, 8>", "zero_push_nvls", ["STG", "LDG"])],
    "nvls_allreduce.o": [(r".", "nvls_allreduce", ["LDGMC", "STG"])],
    "mt_optim.o": [(r"AdamOp<false, false>", "mt_adam", ["LDG", "STG"]), (r"SgdOp", "mt_sgd", ["LDG", "STG"]),
                   (r"NovoGrad", "mt_novograd", ["LDG", "STG"]), (r"Adagrad", "mt_adagrad", ["LDG", "STG"]),
                   (r"update_scale_hysteresis", "update_scale_hysteresis", ["LDG", "STG"])],
    "mt_basic.o": [(r"L2Norm", "mt_l2norm", ["LDG"]), (r"ScaleOp", "mt_scale", ["LDG", "STG"]), (r"Axpby", "mt_axpby", ["LDG", "STG"])],
    "mt_lamb_dist.o": [(r"LambStage1Op<false, 0>", "mt_lamb_stage1", ["LDG", "STG"]), (r"LambStage2Op<false", "mt_lamb_stage2", ["LDG", "STG"]),
                       (r"DistAdamOp<true>", "mt_dist_adam", ["LDG", "STG"])],
    "layer_norm_fwd.o": [(r"ln_fwd_vec<4, __nv_bfloat16, __nv_bfloat16, false, false>", "layer_norm_fwd_bf16", ["LDG", "STG"])],
    "layer_norm_bwd.o": [(r"ln_bwd_vec<2, __nv_bfloat16, __nv_bfloat16, false, false>", "layer_norm_bwd_bf16", ["LDG", "STG"])],
    "group_norm_stream.o": [(r"gns_stats<__nv_bfloat16, 8, false, false>", "group_norm_stream_stats", ["LDG", "RED"])],
    "softmax.o": [(r"__nv_bfloat16", "softmax_bf16", ["MUFU.EX2", "LDG", "STG"])],
    "xentropy.o": [(r"__nv_bfloat16", "xentropy_bf16", ["MUFU", "LDG", "STG"])],
    "rope.o": [(r"__nv_bfloat16", "rope_bf16", ["LDG", "STG"])],
    "syncbn.o": [(r"syncbn_kernel<__nv_bfloat16, false, true, false>", "syncbn_fwd_nhwc", ["LDG", "STG"]),
                 (r"syncbn_kernel<__nv_bfloat16, true, true, false>", "syncbn_bwd_nhwc", ["LDG", "STG"])],
    "group_norm.o": [(r"__nv_bfloat16", "group_norm_bf16", ["LDG", "STG"])],
    "group_norm_small.o": [(r"__nv_bfloat16", "group_norm_cluster_bf16", ["LDG", "STG", "UCGABAR"])],
    "conv_epilogue.o": [(r"conv_epi_bwd<__nv_bfloat16, 8>", "conv_epilogue_bwd", ["LDG", "STG"]), (r"conv_epi_fwd<__nv_bfloat16, 8>", "conv_epilogue_fwd", ["LDG", "STG"])],
    "halo_exchange.o": [(r".", "halo_exchange", ["LDG", "STG"])],
    "fp8_quant.o": [(r".", "fp8_quant", ["LDG", "STG"])],
    "transducer.o": [(r"__nv_bfloat16|float", "transducer", ["LDG", "STG"])],
    "contrib_ops.o": [(r"focal", "focal_loss", ["LDG", "STG"]), (r"index_mul", "index_mul_2d", ["LDG", "STG"])],
    "perm_search.o": [(r".", "perm_search", ["LDG", "LDS"])],
}
INSTR = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def functions(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, timeout=600).stdout
    cur, funcs = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = INSTR.match(line)
        if m and cur is not None:
            funcs[cur].append((int(m.group(1), 16), m.group(2)))
    return funcs


def loops(ins):
    """(start_index, end_index) of every backward branch, innermost first."""
    addr_to_idx = {a: i for i, (a, _) in enumerate(ins)}
    out = []
    for i, (a, text) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:[!\w]+,\s*)?`?\(?\.?L?_?x?_?\w*\)?", text)
        m2 = re.search(r"BRA.*?0x([0-9a-f]+)", text)
        if "BRA" in text and m2:
            tgt = int(m2.group(1), 16)
            if tgt < a and tgt in addr_to_idx:
                out.append((addr_to_idx[tgt], i))
    out.sort(key=lambda r: r[1] - r[0])
    return out


def usage(obj):
    res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            out[cur] = m.groups()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    index = ["# SASS listings of the hot loops (generated by benchmarks/sass_listing.py from build/obj/*.o, sm_100a)", "",
             "| listing | kernel | registers | stack | markers found |", "|---|---|---|---|---|"]
    for objname, sels in SELECT.items():
        obj = os.path.join(OBJ, objname)
        if not os.path.exists(obj):
            continue
        funcs = functions(obj)
        dm = demangle(list(funcs))
        use = usage(obj)
        for rx, tag, markers in sels:
            cands = [f for f in funcs if re.search(rx, dm.get(f, f)) and len(funcs[f]) > 20]
            if not cands:
                continue
            f = max(cands, key=lambda n: len(funcs[n]))
            ins = funcs[f]
            lp = loops(ins)
            chosen, counts = [], {}
            for mk in markers:
                idxs = [i for i, (_, t) in enumerate(ins) if mk in t]
                counts[mk] = len(idxs)
                if not idxs:
                    continue
                inner = [r for r in lp if any(r[0] <= i <= r[1] for i in idxs)]
                if inner:
                    best = max(inner[:6], key=lambda r: sum(r[0] <= i <= r[1] for i in idxs) / (r[1] - r[0] + 8))
                    if best not in chosen:
                        chosen.append(best)
                else:   # straight-line use (e.g. an unrolled epilogue): a window around the first occurrences
                    w = (max(0, idxs[0] - 6), min(len(ins) - 1, idxs[min(len(idxs) - 1, 15)] + 6))
                    chosen.append(w)
            reg = use.get(f, ("?", "?", "?"))
            lines = [f"// {dm.get(f, f)}", f"// registers {reg[0]}, stack {reg[1]} B, static shared {reg[2]} B, {len(ins)} instructions",
                     f"// marker counts in the whole function: {counts}", ""]
            for k, (a, b) in enumerate(sorted(set(chosen))):
                b = min(b, a + 220)
                lines.append(f"// ---- loop / region {k}: {ins[a][0]:#06x} .. {ins[b][0]:#06x} ({b - a + 1} instructions)")
                lines += [f"/*{ad:04x}*/  {t} ;" for ad, t in ins[a:b + 1]]
                lines.append("")
            open(os.path.join(OUT, tag + ".sass"), "w").write("\n".join(lines))
            index.append(f"| [{tag}.sass](sass/{tag}.sass) | `{dm.get(f, f)[:110]}` | {reg[0]} | {reg[1]} | " +
                         ", ".join(f"{k} x{v}" for k, v in counts.items() if v) + " |")
    open(os.path.join(ROOT, "profiles", "sass_listings.md"), "w").write("\n".join(index) + "\n")
    print("\n".join(index[-50:]))


if __name__ == "__main__":
    main()
