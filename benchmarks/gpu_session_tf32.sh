#!/bin/bash
# GEMM family check after the tf32 path / 3-D MN-major boxes: tests, bf16 table (3-D boxes on and off), tf32 table vs cuBLAS TF32.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_modules.py -x -q 2>&1 | tail -6
timeout 300 python benchmarks/bench_gemm.py > gpurun_out/bench_gemm_3d.log 2>&1; tail -12 gpurun_out/bench_gemm_3d.log
cp profiles/results/bench_gemm.json gpurun_out/bench_gemm_3d.json 2>/dev/null
APEX_B200_GEMM_NO3D=1 timeout 300 python benchmarks/bench_gemm.py > gpurun_out/bench_gemm_no3d.log 2>&1; tail -12 gpurun_out/bench_gemm_no3d.log
for mode in 3d no3d; do
if [ $mode = no3d ]; then export APEX_B200_GEMM_NO3D=1; fi
PYTHONPATH=. timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/bench_tf32_$mode.txt
import torch, json, os
from apex_b200.ops import gemm as G
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n
rows = []
for (M, N, K) in [(8192, 4096, 4096), (8192, 16384, 4096), (4096, 4096, 4096), (1536, 3072, 1024)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); dy = torch.randn(M, N, device=dev)
    for name, ours, ref in [("fwd", lambda: G.linear_fwd(x, w), lambda: x @ w.t()),
                            ("dgrad", lambda: G.linear_dgrad(dy, w), lambda: dy @ w),
                            ("wgrad", lambda: G.linear_wgrad(dy, x), lambda: dy.t() @ x)]:
        a, b = t(ours), t(ref)
        fl = 2.0 * M * N * K
        rows.append({"shape": [M, N, K], "op": name, "ours_ms": round(a, 4), "cublas_tf32_ms": round(b, 4), "ours_tflops": round(fl / a / 1e9, 1), "ratio": round(b / a, 3)})
        print(rows[-1])
json.dump(rows, open("gpurun_out/bench_tf32_%s.json" % ("no3d" if os.environ.get("APEX_B200_GEMM_NO3D") else "3d"), "w"), indent=1)
PY
done
