"""FFN-block GEMM benchmark (BASELINE.md compute-bound target): tokens M=8192, hidden 4096, ffn 16384, bf16.
Our tcgen05 kernel vs cuBLAS (torch.matmul) vs the reference fused_dense_cuda (cuBLASLt) when baseline/_ref is present.
CUDA-event timed; operands (64-256 MB) are re-read from HBM/L2 as in real training; L2 flushed between iterations."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from apex_b200.ops import gemm as G  # noqa: E402
from apex_b200.utils.timing import measured_peaks, time_fn  # noqa: E402


def tflops(m, n, k, ms):
    return 2.0 * m * n * k / ms / 1e9


def main():
    out = []
    pk = measured_peaks()
    dev = "cuda"
    M, H, FF = 8192, 4096, 16384
    x = torch.randn(M, H, device=dev, dtype=torch.bfloat16)
    w1 = torch.randn(FF, H, device=dev, dtype=torch.bfloat16) * 0.02
    b1 = torch.randn(FF, device=dev, dtype=torch.bfloat16)
    w2 = torch.randn(H, FF, device=dev, dtype=torch.bfloat16) * 0.02
    dy1 = torch.randn(M, FF, device=dev, dtype=torch.bfloat16)
    aux = torch.empty(M, FF, device=dev, dtype=torch.bfloat16)

    cases = [
        ("fwd  x@W1^T        (TN)", lambda: G.gemm(x, w1), lambda: torch.matmul(x, w1.t()), (M, FF, H)),
        ("fwd  +bias+GELU+aux    ", lambda: G.gemm(x, w1, epi=G.EPI_BIAS_GELU, bias=b1, aux=aux),
         lambda: torch.nn.functional.gelu(torch.addmm(b1, x, w1.t())), (M, FF, H)),
        ("dgrad dy@W1        (NN)", lambda: G.gemm(dy1, w1, b_mn=True), lambda: torch.matmul(dy1, w1), (M, H, FF)),
        ("wgrad dy^T@x       (NT)", lambda: G.gemm(dy1, x, a_mn=True, b_mn=True), lambda: torch.matmul(dy1.t(), x), (FF, H, M)),
        ("square 8192^3          ", None, None, (8192, 8192, 8192)),
    ]
    a8 = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    b8 = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    for name, ours, cublas, (m, n, k) in cases:
        if ours is None:
            ours, cublas = (lambda: G.gemm(a8, b8)), (lambda: torch.matmul(a8, b8.t()))
        assert ours() is not None
        t_o, t_o_min = time_fn(ours, warmup=3, iters=10)
        t_c, t_c_min = time_fn(cublas, warmup=3, iters=10)
        r = {"case": name.strip(), "M": m, "N": n, "K": k, "ours_ms": t_o, "cublas_ms": t_c, "ours_tflops": tflops(m, n, k, t_o),
             "cublas_tflops": tflops(m, n, k, t_c), "ours_frac_of_measured_peak": tflops(m, n, k, t_o) / pk["bf16_tflops"],
             "ours_vs_cublas": t_c / t_o}
        out.append(r)
        print(json.dumps(r))

    # whole FusedDenseGeluDense block fwd+bwd, ours vs reference extension
    from apex_b200.fused_dense import FusedDenseGeluDense
    blk = FusedDenseGeluDense(H, FF, H).to(dev, torch.bfloat16)
    xin = torch.randn(M, H, device=dev, dtype=torch.bfloat16, requires_grad=True)
    dout = torch.randn(M, H, device=dev, dtype=torch.bfloat16)

    def fb(mod):
        y = mod(xin)
        y.backward(dout)

    t, _ = time_fn(lambda: fb(blk), warmup=3, iters=10)
    flops = 6.0 * 2 * M * H * FF
    r = {"case": "FusedDenseGeluDense fwd+bwd (ours)", "ms": t, "tflops": flops / t / 1e9, "frac_of_measured_peak": flops / t / 1e9 / pk["bf16_tflops"]}
    out.append(r)
    print(json.dumps(r))
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "apex")):
        sys.path.insert(0, ref_dir)
        try:
            from apex.fused_dense import FusedDenseGeluDense as RefBlk
            rb = RefBlk(H, FF, H).to(dev, torch.bfloat16)
            t, _ = time_fn(lambda: fb(rb), warmup=3, iters=10)
            r = {"case": "FusedDenseGeluDense fwd+bwd (reference, cuBLASLt)", "ms": t, "tflops": flops / t / 1e9}
            out.append(r)
            print(json.dumps(r))
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"case": "reference fused_dense", "unavailable": str(e)[:200]}))
    tm = torch.nn.Sequential(torch.nn.Linear(H, FF), torch.nn.GELU(), torch.nn.Linear(FF, H)).to(dev, torch.bfloat16)
    t, _ = time_fn(lambda: fb(tm), warmup=3, iters=10)
    r = {"case": "torch Linear-GELU-Linear fwd+bwd", "ms": t, "tflops": flops / t / 1e9}
    out.append(r)
    print(json.dumps(r))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"peaks": pk, "results": out}, open(os.path.join(ROOT, "gpurun_out", "bench_gemm.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
