"""fp8 (E4M3) tcgen05 GEMM: TFLOP/s of apex_b200.ops.gemm.gemm_fp8 vs torch._scaled_mm (cuBLASLt fp8) vs our bf16 kernel, on the FFN
shapes of bench_gemm.py. Operands are quantised once outside the timed region (the GEMM itself is what is measured)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apex_b200.ops import gemm as G  # noqa: E402
from apex_b200.utils.timing import time_fn  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    out = []
    for name, M, N, K in (("fwd x@W1^T", 8192, 16384, 4096), ("fwd h@W2^T", 8192, 4096, 16384), ("square", 8192, 8192, 8192)):
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        a8, sa = G.quantize_fp8(a)
        b8, sb = G.quantize_fp8(b)
        alpha = float((sa * sb).item())
        flops = 2.0 * M * N * K
        t8 = time_fn(lambda: G.gemm_fp8(a8, b8, alpha), warmup=3, iters=10)[0]
        t16 = time_fn(lambda: G.gemm(a, b), warmup=3, iters=10)[0]
        row = {"case": f"{name} {M}x{N}x{K}", "ours_fp8_ms": round(t8, 4), "ours_fp8_tflops": round(flops / t8 / 1e9, 1),
               "ours_bf16_ms": round(t16, 4), "ours_bf16_tflops": round(flops / t16 / 1e9, 1)}
        try:
            one = torch.ones((), device=dev)
            f = lambda: torch._scaled_mm(a8, b8.t(), scale_a=one, scale_b=one, out_dtype=torch.bfloat16)  # noqa: E731
            tl = time_fn(f, warmup=3, iters=10)[0]
            row["cublaslt_fp8_ms"], row["cublaslt_fp8_tflops"] = round(tl, 4), round(flops / tl / 1e9, 1)
        except Exception as e:  # noqa: BLE001
            row["cublaslt_fp8"] = f"unavailable: {str(e)[:80]}"
        ref = (a8.float() @ b8.float().t()) * alpha
        got = G.gemm_fp8(a8, b8, alpha, out_dtype=torch.float32)
        row["max_rel_err_vs_dequant_fp32"] = float(((got - ref).abs().max() / ref.abs().max()))
        out.append(row)
        print(json.dumps(row))
        del a, b, a8, b8, ref, got
    # FFN block (BASELINE.md row 3, fp8 variant): fp8 forward GEMMs incl. per-call activation quantisation, 16-bit backward
    from apex_b200.fused_dense import FusedDenseGeluDense, fused_dense_gelu_dense_fp8_function
    M, H, FF = 8192, 4096, 16384
    blk = FusedDenseGeluDense(H, FF, H).to(dev, torch.bfloat16)
    xin = torch.randn(M, H, device=dev, dtype=torch.bfloat16, requires_grad=True)
    dout = torch.randn(M, H, device=dev, dtype=torch.bfloat16)
    f8 = lambda: fused_dense_gelu_dense_fp8_function(xin, blk.weight1, blk.bias1, blk.weight2, blk.bias2)  # noqa: E731
    row = {"case": "FusedDenseGeluDense 8192 x 4096 <-> 16384"}
    with torch.no_grad():
        row["fwd_fp8_ms"] = round(time_fn(f8, warmup=3, iters=10)[0], 4)
        row["fwd_bf16_ms"] = round(time_fn(lambda: blk(xin), warmup=3, iters=10)[0], 4)
    row["fwd_bwd_fp8_ms"] = round(time_fn(lambda: f8().backward(dout), warmup=3, iters=10)[0], 4)
    row["fwd_bwd_bf16_ms"] = round(time_fn(lambda: blk(xin).backward(dout), warmup=3, iters=10)[0], 4)
    f8b = lambda: fused_dense_gelu_dense_fp8_function(xin, blk.weight1, blk.bias1, blk.weight2, blk.bias2, True)  # noqa: E731  fp8 dgrad / wgrad too
    row["fwd_bwd_all_fp8_ms"] = round(time_fn(lambda: f8b().backward(dout), warmup=3, iters=10)[0], 4)
    y8, y16 = f8().float(), blk(xin).float()
    row["fp8_vs_bf16_rel_err"] = float((y8 - y16).norm() / y16.norm())
    out.append(row)
    print(json.dumps(row))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_gemm_fp8.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
