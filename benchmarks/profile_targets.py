"""Small driver for ncu: runs each hot kernel a few times (after warm-up) so `ncu -k regex:... -s N -c M` can capture it.
usage: python benchmarks/profile_targets.py <target>   with target in {dist_adam, gemm, layer_norm, adam, syncbn, softmax, xent, group_norm}"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "adam"
torch.manual_seed(0)

if what == "dist_adam":
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    ps = [torch.nn.Parameter(torch.randn(1 << 28, device=dev, dtype=torch.bfloat16))]
    opt = DistributedFusedAdam(ps, lr=1e-3, weight_decay=0.1)
    opt.zero_grad()
    ps[0].grad.normal_()
    for _ in range(4):
        opt.step()
elif what == "adam":
    from apex_b200.optimizers import FusedAdam
    ps = [torch.nn.Parameter(torch.randn(1 << 28, device=dev))]
    ps[0].grad = torch.randn_like(ps[0])
    opt = FusedAdam(ps, lr=1e-3)
    for _ in range(4):
        opt.step()
elif what == "gemm":
    from apex_b200.ops import gemm as G
    x = torch.randn(8192, 4096, device=dev, dtype=torch.bfloat16)
    w = torch.randn(16384, 4096, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        G.gemm(x, w)
elif what == "layer_norm":
    from apex_b200.normalization import FusedLayerNorm
    m = FusedLayerNorm(4096).to(dev, torch.bfloat16)
    x = torch.randn(131072, 4096, device=dev, dtype=torch.bfloat16, requires_grad=True)
    for _ in range(3):
        y = m(x)
        y.backward(torch.ones_like(y))
elif what == "syncbn":
    from apex_b200.parallel import SyncBatchNorm
    for (C, HW) in [(64, 112), (512, 7)]:
        bn = SyncBatchNorm(C).to(dev)
        x = torch.randn(64, C, HW, HW, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        for _ in range(3):
            bn(x).backward(torch.ones_like(x))
elif what == "softmax":
    from apex_b200.transformer.functional import scaled_upper_triang_masked_softmax
    x = torch.randn(64, 4096, 4096, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        scaled_upper_triang_masked_softmax(x, 0.5)
elif what == "xent":
    from apex_b200.contrib.xentropy import SoftmaxCrossEntropyLoss
    x = torch.randn(9472, 32320, device=dev, dtype=torch.bfloat16, requires_grad=True)
    lab = torch.randint(0, 32320, (9472,), device=dev)
    for _ in range(3):
        SoftmaxCrossEntropyLoss.apply(x, lab, 0.1, 0, True).sum().backward()
elif what == "group_norm":
    from apex_b200.contrib.group_norm import GroupNorm
    gn = GroupNorm(32, 640, act="silu").to(dev, torch.bfloat16)
    x = torch.randn(8, 640, 64, 64, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    for _ in range(3):
        gn(x).backward(torch.ones_like(x))
torch.cuda.synchronize()
print("done", what)
