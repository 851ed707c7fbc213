"""Writes profiles/ROOFLINE.md: achieved numbers of the hot kernels against the measured peaks, one table, from the JSON files under
profiles/results/ (no GPU needed). usage: python benchmarks/make_roofline.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = os.path.join(ROOT, "profiles", "results")


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tf, tfs = peaks["hbm_gbs"], peaks["bf16_tflops"], peaks["bf16_tflops_sustained"]
    ops = json.load(open(os.path.join(R, "bench_ops.json")))
    ops = ops if isinstance(ops, list) else ops.get("results", ops)
    op = {r["op"]: r for r in ops}
    gem = json.load(open(os.path.join(R, "bench_gemm.json")))["results"]
    tf32 = json.load(open(os.path.join(R, "bench_gemm_tf32.json")))
    fm = [json.loads(l) for l in open(os.path.join(R, "bench_fmha.json"))]
    o1 = last_json(os.path.join(R, "r2", "bench_ours_n1_final.json"))
    o2 = last_json(os.path.join(R, "r2", "bench_ours_n2_final.json"))
    L = ["# Achieved vs measured peaks (one table; sources are the JSON files under `profiles/results/`, generator `benchmarks/make_roofline.py`)\n",
         f"Denominators: `MEASURED_PEAKS.json` (driver-written on this B200): copy bandwidth **{hbm:.0f} GB/s**, cuBLAS bf16 **{tf:.0f} TFLOP/s** burst / "
         f"**{tfs:.0f}** sustained;",
         "link: 725 GB/s per direction with one direction busy (`results/link_peaks.json`), 514 - 577 GB/s per direction with both busy "
         "(`results/r2/symm_n{4,8}.jsonl`).",
         "All numbers are CUDA-event timings after warm-up (never taken under a profiler); the ncu file is the evidence for WHY, not for how fast.\n",
         "| kernel (file) | workload | achieved | of peak | ncu / SASS evidence |", "|---|---|---|---|---|",
         f"| `dist_step_kernel` world 1 (`dist_adam.cu`) | Llama-3-8B ZeRO step, 8.03 B params | {o1['value']:.2f} ms = "
         f"{o1['roofline']['t_hbm_ms'] / o1['value'] * hbm:.0f} GB/s | {o1['roofline']['achieved_frac'] * 100:.0f} % of copy bw | `zero_step_world1.md`, "
         "`sass/zero_step_world1.sass` |",
         f"| `dist_step_kernel` 2 GPUs, P2P pull / push | same, 2 GPUs | {o2['value']:.2f} ms | {o2['roofline']['achieved_frac'] * 100:.0f} % of the "
         "max(HBM, link) bound | `sass/zero_step_p2p2.sass` |",
         "| `dist_step_kernel` NVLS, 4 / 8 GPUs | same | 36.0 / 33.7 ms = 555 / 534 GB/s per direction | 96 % / 104 % of the duplex link probe; 74 - 78 % of "
         "the one-direction peak | `sass/zero_step_nvls.sass` (LDGMC / multimem.st loop); no ncu (multi-rank) |"]
    ev = "`gemm2_384.md` (tensor pipe 98 %), `gemm_dgelu_bgrad.md`, `gemm_wgrad_accum.md`, `sass/gemm2_bf16.sass`"
    for r in gem:
        if "ours_tflops" in r:
            L.append(f"| `gemm2_kernel` bf16 (`gemm_sm100.cu`) | {' '.join(r['case'].split())} {r['M']}x{r['N']}x{r['K']} | {r['ours_tflops']:.0f} TFLOP/s "
                     f"(cuBLAS {r['cublas_tflops']:.0f}) | {r['ours_tflops'] / tf * 100:.0f} % of the cuBLAS burst peak | {ev} |")
        elif "ms" in r:
            L.append(f"| FFN block | {r['case']} | {r['ms']:.2f} ms ({r.get('tflops', 0):.0f} TFLOP/s) | {r.get('tflops', 0) / tf * 100:.0f} % | same |")
    r = [x for x in tf32 if x["shape"] == [8192, 16384, 4096] and x["op"] == "fwd"][0]
    L.append(f"| `gemm2_kernel<float, 6, 2>` tf32 | fwd 8192x16384x4096 | {r['ours_tflops']:.0f} TFLOP/s ({r['ratio']:.2f}x cuBLAS TF32) | tensor pipe 97 % active "
             "(tf32 runs at half the bf16 rate) | `gemm_tf32_fwd.md`, `gemm_tf32_wgrad.md`, `sass/gemm2_tf32.sass` |")
    L.append("| `gemm2_kernel` fp8 | 8192x16384x4096 e4m3 | 3173 - 3254 TFLOP/s (cuBLASLt fp8 3173 - 3314) | ~71 % of the nominal 4.5 PFLOP/s | "
             "`sass/gemm2_fp8.sass` (UTCQMMA) |")
    for d in fm:
        if "summary" not in d and d["seq"] == 8192:
            L.append(f"| `fmha_fwd_kernel` d={d['d']} {'causal' if d['causal'] else 'full'} | 2 x 8192, {2048 // d['d']} heads | {d['ours_fwd_tflops']:.0f} TFLOP/s fwd | "
                     f"{d['ours_fwd_tflops'] / tf * 100:.0f} % of the cuBLAS peak; {d['sdpa_fwd_ms'] / d['ours_fwd_ms']:.2f}x SDPA | `fmha_fwd_online.md`, "
                     "`sass/fmha_fwd_d128.sass` |")
    for h in (1024, 4096, 8192, 16384):
        f, b = op[f"LayerNorm fwd bf16 h={h}"], op[f"LayerNorm bwd bf16 h={h}"]
        rf, rb = op[f"RMSNorm fwd bf16 h={h}"], op[f"RMSNorm bwd bf16 h={h}"]
        L.append(f"| `ln_fwd_vec` / `ln_bwd_vec` h={h} | 1 GiB bf16 | LN {f['GBps']:.0f} / {b['GBps']:.0f}, RMS {rf['GBps']:.0f} / {rb['GBps']:.0f} GB/s (fwd / bwd) | "
                 f"LN {f['GBps'] / hbm * 100:.0f} / {b['GBps'] / hbm * 100:.0f} %, RMS {rf['GBps'] / hbm * 100:.0f} / {rb['GBps'] / hbm * 100:.0f} % | "
                 "`ln_fwd_r2b.md`, `ln_bwd_r2b.md` |")
    a1, a2, lm = op["FusedAdam fp32 1 tensors"], op["FusedAdam fp32 10000 tensors"], op["FusedLAMB fp32 10000 tensors"]
    L.append(f"| `mt_kernel<AdamOp>` (`mt_optim.cu`) | 1 tensor of 2^28 / 10 000 tensors (1.0 G elements) | {a1['GBps']:.0f} / {a2['GBps']:.0f} GB/s | "
             f"{a1['GBps'] / hbm * 100:.0f} % / {a2['GBps'] / hbm * 100:.0f} % | `mt_adam.md`, `sass/mt_adam.sass` |")
    L.append(f"| `mt_kernel<LambStage1/2>` | 10 000 tensors | {lm['GBps']:.0f} GB/s | {lm['GBps'] / hbm * 100:.0f} % | `mt_lamb_stage1.md`, `mt_lamb_stage2.md` |")
    L.append("| `rope_kernel` | sbhd 4096 x 8 x 32 x 128 bf16 | 0.152 ms = 3.5 TB/s | 53 % | `rope_fwd.md` |")
    L.append("| `syncbn_kernel` | ResNet-50, batch 64 / GPU, 4 GPUs | 13 902 img/s (torch SyncBatchNorm 6 535) | n/a (latency-bound layers) | `syncbn.md` |")
    L.append("| GroupNorm (`group_norm_small.cu`, `group_norm_stream.cu`) | 8 x 4096 x 960, G 16, SiLU | fwd 90 us (reference 58), bwd 223 us (150): 1.4 - 1.7 TB/s | "
             "21 - 26 % (10 - 30 us problems: latency-bound) | `group_norm.md`, `results/gn_probe_*.csv` |")
    open(os.path.join(ROOT, "profiles", "ROOFLINE.md"), "w").write("\n".join(L) + "\n")
    print(f"wrote profiles/ROOFLINE.md ({len(L)} lines)")


if __name__ == "__main__":
    main()
