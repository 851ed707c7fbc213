"""Condense `ncu --page raw --csv` exports (gpurun_out/ncu/*.raw.csv) into profiles/<name>.md: the roofline-relevant metrics per
captured kernel. usage: python benchmarks/ncu_summary.py [names...]"""
import csv
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "nvlrx__bytes.sum", "nvltx__bytes.sum", "pcie__read_bytes.sum",
]


def load(path):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    return hdr, units, [r for r in data if len(r) == len(hdr)]


def main():
    names = sys.argv[1:] or sorted(os.path.basename(p)[:-8] for p in glob.glob(os.path.join(ROOT, "gpurun_out/ncu/*.raw.csv")))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for name in names:
        path = os.path.join(ROOT, "gpurun_out/ncu", name + ".raw.csv")
        if not os.path.exists(path):
            continue
        hdr, units, data = load(path)
        col = {h: i for i, h in enumerate(hdr)}
        out = [f"# ncu --set full --clock-control none: {name}", "",
               "Source: `gpurun_out/ncu/%s.raw.csv` (one B200, kernel replay; durations under ncu are NOT benchmark numbers)." % name, ""]
        for r in data:
            out.append(f"## `{r[col['Kernel Name']][:140]}`  grid {r[col.get('Grid Size', 0)]} block {r[col.get('Block Size', 0)]}")
            out.append("")
            out.append("| metric | value | unit |")
            out.append("|---|---|---|")
            for k in KEYS:
                if k in col and r[col[k]] not in ("", "n/a"):
                    out.append(f"| {k} | {r[col[k]]} | {units[col[k]]} |")
            extra = [h for h in hdr if ("stalled" in h and h.endswith("per_issue_active.ratio") and h not in KEYS)]
            top = sorted(((float(r[col[h]].replace(",", "")) if r[col[h]] not in ("", "n/a") else 0.0, h) for h in extra), reverse=True)[:4]
            for v, h in top:
                if v > 0.05:
                    out.append(f"| {h} | {v:.3f} | (other top stall) |")
            out.append("")
        open(os.path.join(ROOT, "profiles", name + ".md"), "w").write("\n".join(out) + "\n")
        print("wrote profiles/%s.md (%d kernels)" % (name, len(data)))


if __name__ == "__main__":
    main()
