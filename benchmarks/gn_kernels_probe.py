"""Runs one GroupNorm forward + backward at a large shape (for `ncu --metrics gpu__time_duration.sum`): which kernels run and how long each takes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1] if len(sys.argv) > 1 else "ours"
HW, C, G, act = int(os.environ.get("GN_HW", 4096)), int(os.environ.get("GN_C", 960)), int(os.environ.get("GN_G", 16)), os.environ.get("GN_ACT", "silu")
side = int(HW ** 0.5)
x = torch.randn(8, C, side, side, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
dy = torch.randn_like(x)
if which == "ours":
    from apex_b200.contrib.group_norm import GroupNorm
else:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref"))
    try:
        import group_norm_cuda  # noqa: F401
    except ImportError:
        import types
        sys.modules["group_norm_cuda"] = types.ModuleType("group_norm_cuda")
    from apex.contrib.group_norm import GroupNorm
m = GroupNorm(G, C, act=act).cuda().bfloat16()
for _ in range(3):
    y = m(x)
    y.backward(dy)
torch.cuda.synchronize()
