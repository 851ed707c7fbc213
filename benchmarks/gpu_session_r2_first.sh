#!/bin/bash
# First GPU session of the next round: everything written after the round-1 GPU budget ran out, cheapest and safest first.
# Build the experimental kernels HERE before calling gpurun (the .so travels with the snapshot):
#     APEX_B200_EXPERIMENTAL=1 python -m apex_b200._build
#     benchmarks/gpuretry.sh /tmp/r2a.log --timeout 1500 -- 'bash benchmarks/gpu_session_r2_first.sh'
# Every kernel wait is bounded (traps after ~2 s / ~10 s), and every step runs under its own timeout.
mkdir -p gpurun_out
export APEX_B200_UNVERIFIED_TESTS=1
echo "== 1. regression: default GPU suite (must stay green with the relinked library)"
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_ext_compat.py --deselect tests/test_gpu_experimental.py 2>&1 | tail -3 | cut -c1-250
echo "== 2. extension-name shims, native file I/O, torchsched streams, > 2^31-element Adam (each test reports separately)"
timeout 900 python -m pytest tests/test_gpu_ext_compat.py -m gpu -q --timeout 300 2>&1 | tail -25 | cut -c1-300 | tee gpurun_out/r2_ext_compat.log
echo "== 3. experimental kernels: attention forward, then backward, then the NVLS all-reduce (needs >= 2 GPUs)"
timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q --timeout 120 -k "fmha_fwd" 2>&1 | tail -25 | cut -c1-300 | tee gpurun_out/r2_fmha_fwd.log
timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q --timeout 120 -k "fmha_bwd" 2>&1 | tail -25 | cut -c1-300 | tee gpurun_out/r2_fmha_bwd.log
timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q --timeout 180 -k "not fmha" 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/r2_nvls.log
echo "== 4. contrib.fmha through the kernels (opt-in route) against the SDPA route"
APEX_B200_FMHA_KERNEL=1 timeout 300 python - <<'PY' 2>&1 | tail -5 | cut -c1-300
import torch
from apex_b200.contrib.fmha.fmha import fmha_varlen
from apex_b200.utils import config
torch.manual_seed(0)
lens = [128, 77, 300, 512]
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
qkv = torch.randn(sum(lens), 3, 16, 64, device="cuda", dtype=torch.float16)
a = fmha_varlen(qkv, cu, max(lens), 0.0, False)
import os; os.environ["APEX_B200_FMHA_KERNEL"] = "0"
b = fmha_varlen(qkv, cu, max(lens), 0.0, False)
print("fmha kernel vs sdpa max abs diff:", (a.float() - b.float()).abs().max().item())
PY
echo "== 5. (run this script under gpurun --gpus 2 or more) symmetric-heap bandwidth / barrier probes"
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 300 python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node "$(nvidia-smi -L | wc -l)" benchmarks/bench_symm.py --mb 256 2>&1 | grep "^{" | tee gpurun_out/r2_symm.jsonl
fi
