"""GroupNorm (+SiLU) NHWC fwd / bwd on the reference's 14 Blackwell (HW, C) shapes (apex/contrib/csrc/group_norm_v2/gn_dispatch_hw_c.hpp:3-62):
ours vs torch.nn.functional.group_norm (+silu) vs the reference module (apex.contrib.group_norm.GroupNorm from baseline/_ref, which
auto-selects its group_norm_v2 Blackwell kernels) when its extensions are built. CUDA events, L2 flushed between iterations."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apex_b200.utils.timing import time_fn  # noqa: E402

SHAPES = [(64, 1280), (64, 2560), (256, 640), (256, 1280), (256, 1920), (256, 2560), (1024, 320), (1024, 640), (1024, 960), (1024, 1280),
          (1024, 1920), (4096, 320), (4096, 640), (4096, 960)]


def main():
    dev = torch.device("cuda:0")
    from apex_b200.contrib.group_norm import GroupNorm as Ours
    ref_cls = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        try:
            import group_norm_cuda  # noqa: F401  (v1 library, 104 MB: not shipped to the GPU box, see .gpurunignore)
        except ImportError:
            import types
            sys.modules["group_norm_cuda"] = types.ModuleType("group_norm_cuda")  # v2 is auto-selected on sm_100; v1 is never called
        from apex.contrib.group_norm import GroupNorm as ref_cls  # noqa: F811
    except Exception as e:  # extension not built
        print(json.dumps({"reference_group_norm": "unavailable", "why": f"{type(e).__name__}: {e}"[:200]}))
    batch = int(os.environ.get("GN_BATCH", 8))
    rows = []
    for G, act in ((32, ""), (16, "silu")):
        for HW, C in SHAPES:
            side = int(HW ** 0.5)
            x = torch.randn(batch, C, side, side, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
            dy = torch.randn_like(x)
            impls = {"ours": Ours(G, C, act=act).to(dev, torch.bfloat16)}
            if ref_cls is not None:
                impls["reference"] = ref_cls(G, C, act=act).to(dev, torch.bfloat16)
            tg = torch.nn.GroupNorm(G, C).to(dev, torch.bfloat16)
            impls["torch"] = (lambda t: (lambda v: torch.nn.functional.silu(t(v)) if act else t(v)))(tg)
            row = {"G": G, "act": act or "none", "N": batch, "HW": HW, "C": C}
            nbytes = x.numel() * 2
            for name, m in impls.items():
                xi = x.detach().clone().requires_grad_(True)
                y = m(xi)
                f = time_fn(lambda: m(xi), warmup=3, iters=10)[0]
                b = time_fn(lambda: y.backward(dy, retain_graph=True), warmup=3, iters=10)[0]
                row[name + "_fwd_us"], row[name + "_bwd_us"] = round(f * 1e3, 1), round(b * 1e3, 1)
                if name == "ours":
                    row["ours_fwd_GBps"], row["ours_bwd_GBps"] = round(2 * nbytes / f / 1e6, 0), round(3 * nbytes / b / 1e6, 0)
            rows.append(row)
            print(json.dumps(row))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_group_norm.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
