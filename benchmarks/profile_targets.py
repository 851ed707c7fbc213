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
elif what == "gemm_tf32":
    from apex_b200.ops import gemm as G
    torch.backends.cuda.matmul.allow_tf32 = True
    x = torch.randn(8192, 4096, device=dev)
    w = torch.randn(16384, 4096, device=dev)
    dyy = torch.randn(8192, 16384, device=dev)
    for _ in range(3):
        G.gemm(x, w)
        G.linear_wgrad(dyy, x)
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
elif what in ("lamb", "sgd", "novograd"):
    from apex_b200.optimizers import FusedLAMB, FusedNovoGrad, FusedSGD
    ps = [torch.nn.Parameter(torch.randn(1 << 27, device=dev))]
    ps[0].grad = torch.randn_like(ps[0])
    opt = {"lamb": lambda: FusedLAMB(ps, lr=1e-3), "sgd": lambda: FusedSGD(ps, lr=1e-3, momentum=0.9),
           "novograd": lambda: FusedNovoGrad(ps, lr=1e-3)}[what]()
    for _ in range(3):
        opt.step()
elif what == "mt_basic":
    from apex_b200.multi_tensor_apply import multi_tensor_applier
    from apex_b200.ops import amp_C
    xs = [torch.randn(1 << 26, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    ys = [torch.empty_like(x) for x in xs]
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    for _ in range(3):
        multi_tensor_applier(amp_C.multi_tensor_l2norm, flag, [xs], False)
        multi_tensor_applier(amp_C.multi_tensor_scale, flag, [xs, ys], 0.5)
        multi_tensor_applier(amp_C.multi_tensor_axpby, flag, [xs, ys, ys], 1.0, 2.0, -1)
        amp_C.update_scale_hysteresis(torch.ones(1, device=dev), torch.zeros(1, dtype=torch.int32, device=dev),
                                      torch.ones(1, dtype=torch.int32, device=dev), torch.zeros(1, device=dev), 2.0, 0.5, 2000, 2)
elif what == "softmax_bwd":
    from apex_b200.transformer.functional import scaled_upper_triang_masked_softmax
    x = torch.randn(32, 4096, 4096, device=dev, dtype=torch.bfloat16, requires_grad=True)
    for _ in range(3):
        y = scaled_upper_triang_masked_softmax(x, 0.5)
        y.backward(torch.ones_like(y))
elif what == "rope":
    from apex_b200.transformer.functional import fused_apply_rotary_pos_emb
    t = torch.randn(4096, 8, 32, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
    freqs = torch.randn(4096, 1, 1, 128, device=dev)
    for _ in range(3):
        y = fused_apply_rotary_pos_emb(t, freqs)
        y.backward(torch.ones_like(y))
elif what == "wgrad":
    from apex_b200.ops import gemm as G
    x = torch.randn(8192, 4096, device=dev, dtype=torch.bfloat16)
    dy = torch.randn(8192, 16384, device=dev, dtype=torch.bfloat16)
    main_grad = torch.zeros(16384, 4096, device=dev, dtype=torch.float32)
    for _ in range(3):
        G.linear_wgrad(dy, x, accum_into=main_grad)
elif what == "dgelu_bgrad":
    from apex_b200.ops import gemm as G
    dy = torch.randn(8192, 4096, device=dev, dtype=torch.bfloat16)
    w2 = torch.randn(4096, 16384, device=dev, dtype=torch.bfloat16)
    aux = torch.randn(8192, 16384, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        G.linear_dgrad(dy, w2, dgelu_aux=aux, want_colsum=True)
elif what == "fmha":
    from apex_b200.contrib.fmha import kernels as K
    b, s, h, d = 8, 2048, 16, 128
    qkv = torch.randn(b * s, 3, h, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    for _ in range(3):
        out = K.FmhaFunc.apply(qkv[:, 0], qkv[:, 1], qkv[:, 2], None, None, None, None, b, True, None)
        out.backward(torch.ones_like(out))
elif what == "conv_epilogue":
    from apex_b200.contrib.conv_bias_relu.conv_bias_relu import fused_conv_epilogue
    x = torch.randn(32, 256, 56, 56, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
    w = torch.randn(256, 256, 3, 3, device=dev, dtype=torch.bfloat16, requires_grad=True)
    sc, bi = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev)
    for _ in range(3):
        y = fused_conv_epilogue(x, w, bias=bi, scale=sc, z=x, stride=1, padding=1)
        y.backward(torch.ones_like(y))
elif what == "dist_lamb":
    from apex_b200.contrib.optimizers import DistributedFusedLAMB
    ps = [torch.nn.Parameter(torch.randn(1 << 26, device=dev, dtype=torch.bfloat16))]
    opt = DistributedFusedLAMB(ps, lr=1e-3, weight_decay=0.01)
    for _ in range(3):
        opt.zero_grad()
        ps[0].grad.normal_()
        opt.step()
torch.cuda.synchronize()
print("done", what)
