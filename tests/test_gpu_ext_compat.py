"""GPU paths of the extension-name shims (apex_b200/ext_compat.py): the reference's raw entry points (argument order, return conventions)
on this library's kernels. Part of the default `pytest -m gpu` run (they were opt-in in round 1, before they had ever run on a B200)."""
import importlib
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    import apex_b200

    apex_b200.install_as_apex()
    return importlib.import_module


def test_softmax_family(cuda_dev, ext):
    x = torch.randn(2, 4, 64, 96, device=cuda_dev, dtype=torch.bfloat16)
    mask = torch.rand(2, 1, 64, 96, device=cuda_dev) < 0.3
    y = ext("scaled_masked_softmax_cuda").forward(x, mask, 0.5)
    ref = torch.softmax((x.float() * 0.5).masked_fill(mask, -10000.0), -1)
    torch.testing.assert_close(y.float(), ref, atol=1e-2, rtol=1e-2)
    dy = torch.randn_like(y)
    dx = ext("scaled_masked_softmax_cuda").backward(dy.clone(), y, 0.5)
    ref_dx = 0.5 * ref * (dy.float() - (dy.float() * ref).sum(-1, keepdim=True))
    torch.testing.assert_close(dx.float(), ref_dx, atol=2e-2, rtol=2e-2)
    t = torch.randn(8, 128, 128, device=cuda_dev, dtype=torch.float16)
    y = ext("scaled_upper_triang_masked_softmax_cuda").forward(t, 1.0)
    causal = torch.triu(torch.ones(128, 128, dtype=torch.bool, device=cuda_dev), 1)
    torch.testing.assert_close(y.float(), torch.softmax(t.float().masked_fill(causal, float("-inf")), -1), atol=1e-3, rtol=1e-3)


def test_layer_norm_entry_points(cuda_dev, ext):
    ln = ext("fused_layer_norm_cuda")
    x = torch.randn(512, 1024, device=cuda_dev, dtype=torch.bfloat16)
    w, b = torch.randn(1024, device=cuda_dev, dtype=torch.bfloat16), torch.randn(1024, device=cuda_dev, dtype=torch.bfloat16)
    y, mean, invvar = ln.forward_affine(x, (1024,), w, b, 1e-5)
    torch.testing.assert_close(y.float(), F.layer_norm(x.float(), (1024,), w.float(), b.float(), 1e-5), atol=5e-2, rtol=2e-2)
    dy = torch.randn_like(y)
    dx, dw, db = ln.backward_affine(dy, mean, invvar, x, (1024,), w, b, 1e-5)
    xr = x.float().requires_grad_()
    wr, br = w.float().requires_grad_(), b.float().requires_grad_()
    F.layer_norm(xr, (1024,), wr, br, 1e-5).backward(dy.float())
    torch.testing.assert_close(dx.float(), xr.grad, atol=5e-2, rtol=5e-2)
    torch.testing.assert_close(dw.float(), wr.grad, atol=0.5, rtol=5e-2)
    torch.testing.assert_close(db.float(), br.grad, atol=0.5, rtol=5e-2)
    y, invvar = ln.rms_forward_affine(x, (1024,), w, 1e-5)
    ref = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-5) * w.float()
    torch.testing.assert_close(y.float(), ref, atol=5e-2, rtol=2e-2)
    z, mu, rs = ext("fast_layer_norm").ln_fwd(x, w, b, 1e-5)
    torch.testing.assert_close(z.float(), F.layer_norm(x.float(), (1024,), w.float(), b.float(), 1e-5), atol=5e-2, rtol=2e-2)


def test_dense_and_mlp(cuda_dev, ext):
    fd, mlp_cuda = ext("fused_dense_cuda"), ext("mlp_cuda")
    bf = dict(device=cuda_dev, dtype=torch.bfloat16)
    x, w1, b1, w2, b2 = (torch.randn(256, 512, **bf), torch.randn(1024, 512, **bf) * 0.05, torch.randn(1024, **bf), torch.randn(256, 1024, **bf) * 0.05,
                         torch.randn(256, **bf))
    o1, o2, gelu_in = fd.linear_gelu_linear_forward(x, w1, b1, w2, b2)
    ref = F.linear(F.gelu(F.linear(x.float(), w1.float(), b1.float())), w2.float(), b2.float())
    torch.testing.assert_close(o2.float(), ref, atol=0.1, rtol=5e-2)
    grads = fd.linear_gelu_linear_backward(x, gelu_in, o1, w1, w2, torch.randn_like(o2))
    assert [tuple(g.shape) for g in grads] == [(256, 512), (1024, 512), (1024,), (256, 1024), (256,)]
    outs = mlp_cuda.forward(1, 1, [x, w1, w2, b1, b2])
    ref = F.relu(F.linear(F.relu(F.linear(x.float(), w1.float(), b1.float())), w2.float(), b2.float()))
    torch.testing.assert_close(outs[0].float(), ref, atol=0.1, rtol=5e-2)
    grads = mlp_cuda.backward(1, 1, torch.randn_like(outs[0]), outs, [x, w1, w2, b1, b2])
    assert len(grads) == 5 and grads[0].shape == x.shape


def test_xentropy_focal_index_mul(cuda_dev, ext):
    xe = ext("xentropy_cuda")
    lg, lab = torch.randn(64, 1000, device=cuda_dev), torch.randint(0, 1000, (64,), device=cuda_dev)
    losses, mlse = xe.forward(lg, lab, 0.1, True)
    torch.testing.assert_close(losses, F.cross_entropy(lg, lab, label_smoothing=0.1, reduction="none"), atol=1e-4, rtol=1e-4)
    g = xe.backward(torch.ones(64, device=cuda_dev), lg, mlse, lab, 0.1)
    lr = lg.clone().requires_grad_()
    F.cross_entropy(lr, lab, label_smoothing=0.1, reduction="sum").backward()
    torch.testing.assert_close(g, lr.grad, atol=1e-4, rtol=1e-4)

    from apex_b200.contrib.focal_loss.focal_loss import _ref as focal_ref

    fl = ext("focal_loss_cuda")
    co = torch.randn(4, 100, 80, device=cuda_dev)
    tg = torch.randint(-2, 80, (4, 100), device=cuda_dev)
    npos = torch.tensor([37.0], device=cuda_dev)
    loss, pgrad = fl.forward(co, tg, npos, 80, 0.25, 2.0, 0.0)
    cr = co.clone().requires_grad_()
    ref = focal_ref(cr, tg, npos, 80, 0.25, 2.0, 0.0)
    torch.testing.assert_close(loss.reshape(()), ref.detach(), atol=1e-4, rtol=1e-4)
    ref.backward()
    torch.testing.assert_close(fl.backward(torch.ones((), device=cuda_dev), pgrad, npos), cr.grad, atol=1e-4, rtol=1e-4)

    im = ext("fused_index_mul_2d")
    in1, in2 = torch.randn(50, 64, device=cuda_dev), torch.randn(300, 64, device=cuda_dev)
    idx = torch.randint(0, 50, (300,), device=cuda_dev)
    out = torch.empty_like(in2)
    im.float_forward(out, in1, in2, idx)
    torch.testing.assert_close(out, in1[idx] * in2)
    g1, g2, go = torch.zeros_like(in1), torch.empty_like(in2), torch.randn_like(in2)
    im.float_backward(g1, g2, go, in1, in2, idx)
    torch.testing.assert_close(g2, go * in1[idx])
    torch.testing.assert_close(g1, torch.zeros_like(in1).index_add_(0, idx, go * in2), atol=1e-4, rtol=1e-4)


def test_distopt_adam_entry_points(cuda_dev, ext):
    da = ext("distributed_adam_cuda")
    p0 = torch.randn(100_000, device=cuda_dev)
    pr = p0.clone().requires_grad_()
    opt = torch.optim.AdamW([pr], lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.1)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    pc, mc, vc = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    noop, one = torch.zeros(1, dtype=torch.int32, device=cuda_dev), torch.ones(1, device=cuda_dev)
    lr_t, step_t = torch.tensor([1e-2], device=cuda_dev), torch.zeros(1, dtype=torch.int32, device=cuda_dev)
    for step in range(1, 4):
        gr = torch.randn_like(p0)
        pr.grad = gr.clone()
        opt.step()
        da.multi_tensor_fused_adam(65536, noop, [[p], [m], [v], [gr.clone()], [p]], one, 1e-2, 0.9, 0.99, 1e-8, step, 1, 1, 0.1)
        step_t += 1
        da.multi_tensor_fused_adam_capturable(65536, noop, [[pc], [mc], [vc], [gr.clone()], [pc]], one, lr_t, 0.9, 0.99, 1e-8, step_t, 1, 1, 0.1)
    torch.testing.assert_close(p, pr.detach(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(pc, pr.detach(), atol=1e-5, rtol=1e-5)


def test_fused_adam_large_tensor_64bit_offsets(cuda_dev):
    """> 2^31 elements in ONE tensor (reference tests/L0/run_optimizers/test_adam.py testLargeTensor): element offsets must be 64-bit."""
    from apex_b200.optimizers import FusedAdam

    n = (1 << 31) + 4096 + 3
    p = torch.nn.Parameter(torch.zeros(n, device=cuda_dev, dtype=torch.bfloat16))
    p.grad = torch.zeros(n, device=cuda_dev, dtype=torch.bfloat16)
    probe = torch.tensor([0, 12345, (1 << 31) - 1, 1 << 31, n - 1], device=cuda_dev)
    p.grad[probe] = torch.tensor([1.0, -2.0, 3.0, -4.0, 5.0], device=cuda_dev, dtype=torch.bfloat16)
    opt = FusedAdam([p], lr=0.5, weight_decay=0.0)
    opt.step()
    torch.cuda.synchronize()
    got = p.detach()[probe].float()
    torch.testing.assert_close(got, torch.tensor([-0.5, 0.5, -0.5, 0.5, -0.5], device=cuda_dev), atol=1e-2, rtol=1e-2)  # first step: -lr * sign(g)
    assert float(p.detach()[(1 << 31) + 1]) == 0.0


def test_gds_file_device_roundtrip_multi_chunk(cuda_dev, tmp_path):
    """Device tensors larger than two staging chunks (csrc/file_io.cpp: 32 MB each) exercise the double-buffered pipeline both ways."""
    from apex_b200.contrib.gpu_direct_storage import GDSFile

    a = torch.randn(25_000_001, device=cuda_dev)                     # ~100 MB, not a multiple of the chunk size
    b = torch.randn(3, 5, device=cuda_dev, dtype=torch.bfloat16)
    path = str(tmp_path / "blob.bin")
    with GDSFile(path, "w") as f:
        f.save_data(a)
        f.save_data(b)
    a2, b2 = torch.empty_like(a), torch.empty_like(b)
    with GDSFile(path, "r") as f:
        f.load_data(a2)
        f.load_data(b2)
    assert torch.equal(a, a2) and torch.equal(b, b2)


def test_torchsched_multi_stream_and_cuda_graph(cuda_dev):
    import torch.nn as nn

    from apex_b200.contrib import torchsched as ts

    class Branchy(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c, self.d = nn.Linear(1024, 4096), nn.Linear(4096, 1024), nn.Linear(1024, 64), nn.Linear(1024, 64)
            self.ln = nn.LayerNorm(1024)

        def forward(self, x):
            h = self.b(F.gelu(self.a(x)))
            return self.ln(h + x).sum(-1, keepdim=True) + self.c(x).relu() + torch.tanh(self.d(x))

    torch.manual_seed(0)
    m = Branchy().to(cuda_dev)
    x = torch.randn(512, 1024, device=cuda_dev, requires_grad=True)
    graphs = []

    def backend(gm, example_inputs, **kw):
        graphs.append(ts._backend(gm, example_inputs, **kw))
        return graphs[-1]

    cm = torch.compile(m, backend=backend)
    for _ in range(3):
        out, ref = cm(x), m(x)
        (g,) = torch.autograd.grad(out.sum(), x)
        (g_ref,) = torch.autograd.grad(ref.sum(), x)
        torch.testing.assert_close(out, ref, atol=2e-3, rtol=2e-3)
        torch.testing.assert_close(g, g_ref, atol=2e-3, rtol=2e-3)
    assert graphs[0].plan.streams_used == 2
    with torch.no_grad():
        cg = torch.compile(m, backend=lambda gm, ex: ts._backend(gm, ex, cuda_graph=True))
        for _ in range(3):
            xi = torch.randn(512, 1024, device=cuda_dev)
            torch.testing.assert_close(cg(xi), m(xi), atol=2e-3, rtol=2e-3)


def test_norm_modules_under_torch_compile_use_the_custom_ops(cuda_dev):
    """fullgraph torch.compile of a model with FusedLayerNorm / FusedRMSNorm / GroupNorm on CUDA: the apex_b200:: custom ops keep the fused
    kernels in the graph (no graph break, no fallback), values and gradients equal eager."""
    from apex_b200.contrib.group_norm import GroupNorm
    from apex_b200.normalization import FusedLayerNorm, FusedRMSNorm
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(256, 256), FusedLayerNorm(256), torch.nn.Tanh(), FusedRMSNorm(256)).to(cuda_dev, torch.bfloat16)
    x = torch.randn(64, 256, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    compiled = torch.compile(net, backend="aot_eager", fullgraph=True)
    out, want = compiled(x), net(x)
    torch.testing.assert_close(out.float(), want.float(), atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(torch.autograd.grad(out.float().sum(), x)[0].float(), torch.autograd.grad(want.float().sum(), x)[0].float(), atol=5e-2, rtol=5e-2)
    gn = torch.nn.Sequential(torch.nn.Conv2d(64, 64, 1), GroupNorm(8, 64, act="silu")).to(cuda_dev, torch.bfloat16).to(memory_format=torch.channels_last)
    img = torch.randn(4, 64, 16, 16, device=cuda_dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
    cgn = torch.compile(gn, backend="aot_eager", fullgraph=True)
    out, want = cgn(img), gn(img)
    torch.testing.assert_close(out.float(), want.float(), atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(torch.autograd.grad(out.float().sum(), img)[0].float(), torch.autograd.grad(want.float().sum(), img)[0].float(), atol=5e-2, rtol=5e-2)


def test_contrib_raw_extension_names_gpu(cuda_dev, ext):
    """fused_conv_bias_relu / group_norm_cuda / transducer_*_cuda raw entry points on the GPU kernels."""
    cb, gn, tj, tl = ext("fused_conv_bias_relu"), ext("group_norm_cuda"), ext("transducer_joint_cuda"), ext("transducer_loss_cuda")
    torch.manual_seed(0)
    x = torch.randn(4, 32, 10, 10, device=cuda_dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 32, 3, 3, device=cuda_dev, dtype=torch.float16) * 0.1).contiguous(memory_format=torch.channels_last)
    b = torch.randn(1, 64, 1, 1, device=cuda_dev, dtype=torch.float16)
    out = cb.forward([x, w, b], 1, 1)[0]
    ref = torch.relu(F.conv2d(x.float(), w.float(), b.float().reshape(-1), 1, 1))
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    dy = torch.randn_like(out)
    dx, dw, db = cb.backward([x, w, out, dy], 1, 1)
    g = dy.float() * (out > 0).float()
    torch.testing.assert_close(db.float().reshape(-1), g.sum((0, 2, 3)), atol=0.5, rtol=2e-2)
    assert dx.shape == x.shape and dw.shape == w.shape

    xg = torch.randn(4, 64, 16, 16, device=cuda_dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wg, bg = torch.randn(64, device=cuda_dev, dtype=torch.bfloat16), torch.randn(64, device=cuda_dev, dtype=torch.bfloat16)
    y, sums = gn.forward(xg, 16, wg, bg, 1e-5, 1, True)
    torch.testing.assert_close(y.float(), F.silu(F.group_norm(xg.float(), 16, wg.float(), bg.float(), 1e-5)), atol=5e-2, rtol=5e-2)
    dxg, dwg, dbg = gn.backward(torch.ones_like(y), sums, xg, 16, wg, bg, 1e-5, 1, True)
    assert dxg.shape == xg.shape and dwg.dtype == wg.dtype and dbg.shape == bg.shape

    B, T, U, H, V = 3, 11, 6, 64, 29
    f, g2 = torch.randn(B, T, H, device=cuda_dev), torch.randn(B, U, H, device=cuda_dev)
    f_len = torch.tensor([11, 7, 9], device=cuda_dev, dtype=torch.int32)
    g_len = torch.tensor([6, 4, 5], device=cuda_dev, dtype=torch.int32)
    h = tj.forward(f, g2, f_len, g_len, torch.empty(0, device=cuda_dev), 0, 1, False, True, False, 0.0, 4)
    valid = (torch.arange(T, device=cuda_dev).view(1, T, 1) < f_len.view(B, 1, 1)) & (torch.arange(U, device=cuda_dev).view(1, 1, U) < g_len.view(B, 1, 1))
    torch.testing.assert_close(h[0], torch.relu(f.unsqueeze(2) + g2.unsqueeze(1)) * valid.unsqueeze(-1))
    df, dg = tj.backward([torch.ones_like(h[0]), h[1]], f_len, g_len, torch.empty(0, device=cuda_dev), T, U, False, 1.0)
    torch.testing.assert_close(df, ((f.unsqueeze(2) + g2.unsqueeze(1) > 0) & valid.unsqueeze(-1)).float().sum(2))

    from apex_b200.contrib.transducer.transducer import _TorchTransducerLoss
    logits = torch.randn(B, T, U, V, device=cuda_dev)
    label = torch.randint(1, V, (B, U - 1), device=cuda_dev, dtype=torch.int32)
    y_len = g_len - 1
    lp = torch.log_softmax(logits, -1)
    alpha, beta, loss = tl.forward(lp, label, f_len, y_len, torch.empty(0, device=cuda_dev), T, 0, 1, False)
    lr = logits.detach().clone().requires_grad_(True)
    ref_loss = _TorchTransducerLoss.forward(_TorchTransducerLoss(), lr, label, f_len, y_len, 0)
    torch.testing.assert_close(loss, ref_loss.detach(), atol=1e-3, rtol=1e-4)
    ref_loss.sum().backward()
    dxl = tl.backward(lp, torch.ones(B, device=cuda_dev), alpha, beta, f_len, y_len, label, torch.empty(0, device=cuda_dev), T, 0, 1, True, False)
    torch.testing.assert_close(dxl, lr.grad, atol=1e-4, rtol=1e-3)


def test_fmhalib_raw_entry_points(cuda_dev, ext):
    """fmhalib.fwd / bwd (packed qkv, cu_seqlens) against autograd through contrib.fmha; S_dmask is an opaque state tensor here."""
    mha = ext("fmhalib")
    from apex_b200.contrib.fmha.fmha import fmha_varlen
    torch.manual_seed(0)
    h, d = 4, 64
    lens = [37, 128, 5, 200]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=cuda_dev)
    qkv = torch.randn(sum(lens), 3, h, d, device=cuda_dev, dtype=torch.float16, requires_grad=True)
    ctx, state = mha.fwd(qkv.detach(), cu, 0.0, max(lens), True, False, False, None)
    ref = fmha_varlen(qkv, cu, max(lens), 0.0, True)
    torch.testing.assert_close(ctx, ref.detach())
    dout = torch.randn_like(ctx)
    ref.backward(dout)
    dqkv, dp = mha.bwd(dout, qkv.detach(), state, cu, 0.0, max(lens), False)
    torch.testing.assert_close(dqkv, qkv.grad)
    # dropout: the backward regenerates the forward's mask from the counters carried in the state tensor
    ctx2, state2 = mha.fwd_nl(qkv.detach(), cu, 0.2, max(lens), True, True, False, None)
    dq2, _, _ = mha.bwd_nl(dout, qkv.detach(), state2, cu, 0.2, max(lens), False)
    assert torch.isfinite(dq2).all() and (ctx2 - ctx).abs().max() > 0
    dq3, _, _ = mha.bwd_nl(dout, qkv.detach(), state2, cu, 0.2, max(lens), False)
    torch.testing.assert_close(dq2, dq3)          # deterministic
