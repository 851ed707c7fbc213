"""The only throughput numbers the reference publishes: 2:4 channel-permutation search on a 64-column x 128-row
random matrix, V100, apex/contrib/sparsity/permutation_tests/README.md:62-97. Same strategies, wall-clock seconds (host loop +
kernels, like the reference's `duration` column), efficacy = share of the gap between default 2:4 and row-pruning that is recovered."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apex_b200.contrib.sparsity import permutation_search as P  # noqa: E402

PUBLISHED_V100_S = {"channel_swap,0": 0.214, "channel_swap,100": 2.249, "channel_swap,1000": 20.248, "optimize_stripe_groups,8,0": 0.013,
                    "optimize_stripe_groups,8,100": 0.152, "optimize_stripe_groups,8,1000": 1.387, "optimize_stripe_groups,12,0": 0.860,
                    "random,1000": 0.116, "random,10000": 1.149, "random,100000": 11.510}


def main():
    dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
    torch.manual_seed(1)
    m = torch.randn(128, 64, device=dev)
    total = float(m.abs().sum())
    base = float(P.sum_after_2_to_4(m))
    rows_opt = float(m.abs().sum(1).topk(64).values.sum())  # "50% rows": the reference's optimistic bound
    # warm-up (module load, candidate tables)
    P.Exhaustive_Search(m, 8)
    P.generate_all_unique_combinations(12, 4)
    out = []
    for name in PUBLISHED_V100_S:
        parts = name.split(",")
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if parts[0] == "channel_swap":
            res, _, perm = P.Channel_Swap(m, escape_attempts=int(parts[1]))
        elif parts[0] == "optimize_stripe_groups":
            res, _, perm = P.Exhaustive_Search(m, stripe_group_size=int(parts[1]), escape_attempts=int(parts[2]))
        else:
            res, _, perm = P.Random_Search(m, num_seeds=int(parts[1]))
        kept = float(P.sum_after_2_to_4(res))
        if dev.type == "cuda":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        row = {"strategy": name, "seconds": round(dt, 4), "published_v100_seconds": PUBLISHED_V100_S[name],
               "speedup_vs_published": round(PUBLISHED_V100_S[name] / dt, 2), "magnitude": round(kept, 3),
               "efficacy": round(P.efficacy(total - rows_opt, total - base, total - kept), 1)}
        out.append(row)
        print(json.dumps(row))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"unpruned": total, "default_2to4": base, "rows": out}, open(os.path.join(ROOT, "gpurun_out", "bench_permutation.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
