"""tcgen05 attention kernels (contrib/fmha/kernels.py) against torch SDPA (flash / cuDNN back-ends) on the same box: forward and
forward+backward, fixed-length batches, head dims 64 / 128, causal and full. CUDA events, L2 flushed between iterations.
usage: python benchmarks/bench_fmha.py > gpurun_out/bench_fmha.json"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apex_b200.contrib.fmha import kernels as K  # noqa: E402
from apex_b200.utils.timing import time_fn  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    rows = []
    for d in (64, 128):
        for causal in (False, True):
            for b, s in ((16, 512), (8, 2048), (2, 8192)):
                h = 2048 // d
                qkv = torch.randn(b * s, 3, h, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
                q4 = qkv.detach().view(b, s, 3, h, d).permute(2, 0, 3, 1, 4).contiguous().requires_grad_()   # [3, b, h, s, d] for SDPA

                def ours_fwd():
                    return K.fmha_fwd(qkv[:, 0], qkv[:, 1], qkv[:, 2], batch=b, causal=causal)

                def ours_fb():
                    o = K.FmhaFunc.apply(qkv[:, 0], qkv[:, 1], qkv[:, 2], None, None, None, None, b, causal, None)
                    o.backward(go)

                def sdpa_fwd():
                    return F.scaled_dot_product_attention(q4[0], q4[1], q4[2], is_causal=causal)

                def sdpa_fb():
                    o = F.scaled_dot_product_attention(q4[0], q4[1], q4[2], is_causal=causal)
                    o.backward(go4)

                go = torch.randn(b * s, h, d, device=dev, dtype=torch.bfloat16)
                go4 = torch.randn(b, h, s, d, device=dev, dtype=torch.bfloat16)
                with torch.no_grad():
                    t_of, _ = time_fn(ours_fwd, 3, 10)
                    t_sf, _ = time_fn(sdpa_fwd, 3, 10)
                t_ob, _ = time_fn(ours_fb, 3, 10)
                t_sb, _ = time_fn(sdpa_fb, 3, 10)
                flops = 4.0 * b * h * s * s * d * (0.5 if causal else 1.0)
                rows.append({"d": d, "causal": causal, "batch": b, "seq": s, "heads": h, "ours_fwd_ms": t_of, "sdpa_fwd_ms": t_sf,
                             "ours_fwd_bwd_ms": t_ob, "sdpa_fwd_bwd_ms": t_sb, "ours_fwd_tflops": flops / t_of / 1e9,
                             "sdpa_fwd_tflops": flops / t_sf / 1e9, "fwd_speedup_vs_sdpa": t_sf / t_of, "fwd_bwd_speedup_vs_sdpa": t_sb / t_ob})
                print(json.dumps(rows[-1]), flush=True)
    print(json.dumps({"summary": {"fwd_geomean_vs_sdpa": float(torch.tensor([r["fwd_speedup_vs_sdpa"] for r in rows]).log().mean().exp()),
                                  "fwd_bwd_geomean_vs_sdpa": float(torch.tensor([r["fwd_bwd_speedup_vs_sdpa"] for r in rows]).log().mean().exp())}}))


if __name__ == "__main__":
    main()
