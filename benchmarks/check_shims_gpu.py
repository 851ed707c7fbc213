"""One-off GPU check of the raw-extension shims that only have CPU tests (cudnn_gbn_lib, permutation_search_cuda)."""
import numpy as np
import torch

from apex_b200 import ext_compat as E

m = E.extension_modules()
g = m["cudnn_gbn_lib"]
torch.manual_seed(0)
dev = "cuda"
x = torch.randn(8, 64, 14, 14, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
w, b = torch.randn(64, device=dev), torch.randn(64, device=dev)
rm, rv, mm, miv = torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.empty(64, device=dev), torch.empty(64, device=dev)
y = g.forward(x, w, b, rm, rv, mm, miv, 0.1, 1e-5, 1, 0, [])
xr, wr, br = x.float().clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
rm2, rv2 = torch.zeros(64, device=dev), torch.ones(64, device=dev)
ref = torch.nn.functional.batch_norm(xr, rm2, rv2, wr, br, True, 0.1, 1e-5)
print("gbn fwd", (y.float() - ref).abs().max().item(), (rm - rm2).abs().max().item(), (rv - rv2).abs().max().item())
dy = torch.randn_like(y)
ref.backward(dy.float())
dx, dw, db = g.backward(x, dy, w, mm, miv, 1e-5, 1, 0, [])
print("gbn bwd", (dx.float() - xr.grad).abs().max().item(), (dw.float() - wr.grad).abs().max().item(), (db.float() - br.grad).abs().max().item())
p = m["permutation_search_cuda"]
M = np.random.default_rng(0).standard_normal((64, 32)).astype(np.float32)
out = np.zeros(1, dtype=np.float32)
p.sum_after_2_to_4(M.flatten(), 64, 32, 0, 32, 4, 16, out)
ref_s = sum(np.sort(np.abs(M[r, c:c + 4]))[2:].sum() for r in range(64) for c in range(0, 32, 4))
print("perm sum", out[0], ref_s)
