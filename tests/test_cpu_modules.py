"""CPU paths of the module layer: every apex-compatible module must run (forward + backward) on CPU tensors and agree with the plain
PyTorch formulation — the same oracles the GPU tests use for the kernels."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F


def test_mlp_and_fused_dense():
    from apex_b200.fused_dense import FusedDense, FusedDenseGeluDense
    from apex_b200.mlp import MLP
    torch.manual_seed(0)
    x = torch.randn(4, 16, requires_grad=True)
    m = MLP([16, 32, 8])
    ref = x
    for w, b in zip(m.weights, m.biases):  # like the reference, the activation follows EVERY layer (tests/L0/run_mlp/test_mlp.py:33-40)
        ref = torch.relu(F.linear(ref, w, b))
    torch.testing.assert_close(m(x), ref)
    a, b2 = FusedDense(16, 8), FusedDenseGeluDense(16, 32, 8)
    torch.testing.assert_close(a(x), F.linear(x, a.weight, a.bias))
    torch.testing.assert_close(b2(x), F.linear(F.gelu(F.linear(x, b2.weight1, b2.bias1)), b2.weight2, b2.bias2))
    (a(x).sum() + b2(x).sum()).backward()
    assert x.grad is not None and b2.weight1.grad is not None


def test_xentropy_focal_index_mul_clip():
    from apex_b200.contrib.clip_grad import clip_grad_norm_
    from apex_b200.contrib.focal_loss import focal_loss
    from apex_b200.contrib.index_mul_2d import index_mul_2d
    from apex_b200.contrib.xentropy import SoftmaxCrossEntropyLoss
    torch.manual_seed(0)
    x = torch.randn(6, 10, requires_grad=True)
    lab = torch.randint(1, 10, (6,))
    loss = SoftmaxCrossEntropyLoss.apply(x, lab, 0.1, 0, True)
    torch.testing.assert_close(loss, F.cross_entropy(x, lab, label_smoothing=0.1, reduction="none"))
    loss.sum().backward()
    a, b = torch.randn(5, 4, requires_grad=True), torch.randn(7, 4, requires_grad=True)
    idx = torch.randint(0, 5, (7,))
    torch.testing.assert_close(index_mul_2d(a, b, idx), a[idx] * b)
    fl = focal_loss(torch.randn(8, 5, requires_grad=True), torch.randint(-2, 4, (8,)), torch.tensor([3.0]), 4, 0.25, 2.0, 0.0)
    assert torch.isfinite(fl).all()
    ps = [nn.Parameter(torch.randn(5)) for _ in range(3)]
    qs = [nn.Parameter(p.detach().clone()) for p in ps]
    for p, q in zip(ps, qs):
        p.grad = torch.randn(5)
        q.grad = p.grad.clone()
    torch.testing.assert_close(clip_grad_norm_(ps, 0.5).reshape(()), torch.nn.utils.clip_grad_norm_(qs, 0.5).reshape(()))
    for p, q in zip(ps, qs):
        torch.testing.assert_close(p.grad, q.grad)


def test_norm_modules_cpu():
    from apex_b200.contrib.group_norm import GroupNorm
    from apex_b200.contrib.layer_norm import FastLayerNorm
    from apex_b200.normalization import FusedLayerNorm, FusedRMSNorm
    torch.manual_seed(0)
    x = torch.randn(4, 32, requires_grad=True)
    ln = FusedLayerNorm(32)
    torch.testing.assert_close(ln(x), F.layer_norm(x, (32,), ln.weight, ln.bias, ln.eps))
    rms = FusedRMSNorm(32)
    torch.testing.assert_close(rms(x), x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + rms.eps) * rms.weight)
    fl = FastLayerNorm(32)
    assert fl(x).shape == x.shape
    g = GroupNorm(4, 16, act="silu")
    img = torch.randn(2, 16, 5, 5).contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(g(img), F.silu(F.group_norm(img, 4, g.weight, g.bias, g.eps)))


def test_batchnorm_family_cpu():
    from apex_b200.contrib.cudnn_gbn import GroupBatchNorm2d
    from apex_b200.contrib.groupbn import BatchNorm2d_NHWC
    from apex_b200.parallel import SyncBatchNorm, convert_syncbn_model
    torch.manual_seed(0)
    m = BatchNorm2d_NHWC(8, fuse_relu=True)
    x, z = torch.randn(2, 4, 4, 8, requires_grad=True), torch.randn(2, 4, 4, 8)
    ref = torch.relu(F.batch_norm(x.permute(0, 3, 1, 2), None, None, m.weight, m.bias, True, 0.1, 1e-5) + z.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    torch.testing.assert_close(m(x, z), ref, atol=1e-5, rtol=1e-5)
    m(x, z).sum().backward()
    assert x.grad.shape == x.shape
    # functional forms with the reference's argument lists (bnp plumbing arguments are ignored): training + eval, gradients vs autograd
    from apex_b200.contrib.groupbn.batch_norm import bn_addrelu_NHWC_impl, bn_NHWC_impl
    w, b = torch.randn(8, requires_grad=True), torch.randn(8, requires_grad=True)
    rm, rv = torch.zeros(8), torch.ones(8)
    xa, za = torch.randn(2, 4, 4, 8, requires_grad=True), torch.randn(2, 4, 4, 8, requires_grad=True)
    scratch = (torch.empty(8), torch.empty(8), torch.zeros(8192, dtype=torch.uint8))
    ipc = (None, None, torch.IntTensor([0]), None, None, 2, 284, 2, 284, False)
    y = bn_NHWC_impl.apply(xa, w, b, rm, rv, *scratch, 0.1, 1e-5, False, True, 1, *ipc)
    rm2, rv2 = torch.zeros(8), torch.ones(8)
    yr = F.batch_norm(xa.permute(0, 3, 1, 2), rm2, rv2, w, b, True, 0.1, 1e-5).permute(0, 2, 3, 1)
    torch.testing.assert_close(y, yr, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(rm, rm2, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(rv, rv2, atol=1e-6, rtol=1e-5)
    gy = torch.randn_like(y)
    for a, r in zip(torch.autograd.grad(y, (xa, w, b), gy), torch.autograd.grad(yr, (xa, w, b), gy)):
        torch.testing.assert_close(a, r, atol=1e-4, rtol=1e-4)
    y = bn_addrelu_NHWC_impl.apply(xa, za, w, b, rm, rv, scratch[0], scratch[1], 1, scratch[2], 0.1, 1e-5, True, 1, *ipc)
    yr = torch.relu(F.batch_norm(xa.permute(0, 3, 1, 2), None, None, w, b, True, 0.1, 1e-5).permute(0, 2, 3, 1) + za)
    torch.testing.assert_close(y, yr, atol=1e-5, rtol=1e-5)
    for a, r in zip(torch.autograd.grad(y, (xa, za, w, b), gy), torch.autograd.grad(yr, (xa, za, w, b), gy)):
        torch.testing.assert_close(a, r, atol=1e-4, rtol=1e-4)
    y = bn_NHWC_impl.apply(xa, w, b, rm, rv, *scratch, 0.1, 1e-5, True, False, 1, *ipc)                  # eval: running statistics + ReLU
    torch.testing.assert_close(y, torch.relu(F.batch_norm(xa.permute(0, 3, 1, 2), rm, rv, w, b, False, 0.0, 1e-5)).permute(0, 2, 3, 1))
    with pytest.raises(RuntimeError, match="bn_group=4"):
        bn_NHWC_impl.apply(xa, w, b, rm, rv, *scratch, 0.1, 1e-5, False, True, 4, *ipc)
    assert GroupBatchNorm2d(8, group_size=1)(torch.randn(2, 8, 4, 4)).shape == (2, 8, 4, 4)
    net = convert_syncbn_model(nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4)))
    assert isinstance(net[1], SyncBatchNorm)
    xi = torch.randn(2, 3, 6, 6)
    torch.testing.assert_close(net(xi), F.batch_norm(net[0](xi), None, None, net[1].weight, net[1].bias, True, 0.1, 1e-5), atol=1e-5, rtol=1e-5)


def test_attention_conv_and_transformer_functional_cpu():
    from apex_b200.contrib.bottleneck import Bottleneck
    from apex_b200.contrib.conv_bias_relu import ConvBias, ConvBiasReLU
    from apex_b200.contrib.multihead_attn import EncdecMultiheadAttn, SelfMultiheadAttn
    from apex_b200.transformer.functional import fused_apply_rotary_pos_emb, scaled_masked_softmax, scaled_upper_triang_masked_softmax
    torch.manual_seed(0)
    m = SelfMultiheadAttn(32, 4, bias=True)
    ref = nn.MultiheadAttention(32, 4, bias=True)
    from apex_b200.contrib.multihead_attn.multihead_attn import blocked_to_packed, packed_to_blocked
    with torch.no_grad():
        nn.init.normal_(m.in_proj_bias)
        # the packed projection is interleaved per head ([head][q|k|v][head_dim] rows, the reference's layout); torch wants [q; k; v] blocks
        ref.in_proj_weight.copy_(packed_to_blocked(m.in_proj_weight, 4))
        ref.in_proj_bias.copy_(packed_to_blocked(m.in_proj_bias, 4))
        ref.out_proj.weight.copy_(m.out_proj_weight)
        ref.out_proj.bias.copy_(m.out_proj_bias)
    assert torch.equal(blocked_to_packed(packed_to_blocked(m.in_proj_weight, 4), 4), m.in_proj_weight)
    x = torch.randn(6, 2, 32)
    torch.testing.assert_close(m(x, x, x, is_training=False)[0], ref(x, x, x, need_weights=False)[0], atol=1e-5, rtol=1e-5)
    # the same layout spelled out as the reference's functions do it: view the projection as [t, b * heads, 3, head_dim]
    lin = F.linear(x, m.in_proj_weight, m.in_proj_bias).view(6, 2 * 4, 3, 8)
    q, k, v = (lin[:, :, i, :].transpose(0, 1) for i in range(3))
    ctx = torch.bmm(torch.softmax(torch.bmm(q, k.transpose(1, 2)) * m.scaling, -1), v).transpose(0, 1).reshape(6, 2, 32)
    torch.testing.assert_close(m(x, is_training=False)[0], F.linear(ctx, m.out_proj_weight, m.out_proj_bias), atol=1e-5, rtol=1e-5)
    sep = SelfMultiheadAttn(32, 4, bias=True, separate_qkv_params=True)
    with torch.no_grad():   # separate q / k / v parameters are ordinary [embed, embed] matrices: same function as the packed module
        for dst, src in zip((sep.q_weight, sep.k_weight, sep.v_weight), packed_to_blocked(m.in_proj_weight, 4).chunk(3)):
            dst.copy_(src)
        for dst, src in zip((sep.q_bias, sep.k_bias, sep.v_bias), packed_to_blocked(m.in_proj_bias, 4).chunk(3)):
            dst.copy_(src)
        sep.out_proj_weight.copy_(m.out_proj_weight)
        sep.out_proj_bias.copy_(m.out_proj_bias)
    torch.testing.assert_close(sep(x, is_training=False)[0], m(x, is_training=False)[0], atol=1e-5, rtol=1e-5)
    enc = EncdecMultiheadAttn(32, 4, bias=True)
    ref2 = nn.MultiheadAttention(32, 4, bias=True)
    mem = torch.randn(5, 2, 32)
    with torch.no_grad():
        nn.init.normal_(enc.in_proj_bias_q)
        nn.init.normal_(enc.in_proj_bias_kv)
        ref2.in_proj_weight.copy_(torch.cat((enc.in_proj_weight_q, packed_to_blocked(enc.in_proj_weight_kv, 4, 2))))
        ref2.in_proj_bias.copy_(torch.cat((enc.in_proj_bias_q, packed_to_blocked(enc.in_proj_bias_kv, 4, 2))))
        ref2.out_proj.weight.copy_(enc.out_proj_weight)
        ref2.out_proj.bias.copy_(enc.out_proj_bias)
    torch.testing.assert_close(enc(x, mem, is_training=False)[0], ref2(x, mem, mem, need_weights=False)[0], atol=1e-5, rtol=1e-5)
    xi, w, b = torch.randn(2, 4, 8, 8), torch.randn(6, 4, 3, 3), torch.randn(1, 6, 1, 1)
    torch.testing.assert_close(ConvBiasReLU(xi, w, b, 1, 1), torch.relu(F.conv2d(xi, w, b.view(-1), 1, 1)))
    torch.testing.assert_close(ConvBias(xi, w, b, 1, 1), F.conv2d(xi, w, b.view(-1), 1, 1))
    assert Bottleneck(16, 8, 32, stride=1)(torch.randn(2, 16, 8, 8)).shape == (2, 32, 8, 8)
    s = torch.randn(2, 5, 5)
    causal = torch.triu(torch.ones(5, 5, dtype=torch.bool), 1)
    torch.testing.assert_close(scaled_upper_triang_masked_softmax(s, 0.5), torch.softmax((s * 0.5).masked_fill(causal, float("-inf")), -1))
    mask = torch.zeros(2, 1, 5, 5, dtype=torch.uint8)
    mask[..., 3:] = 1
    s4 = torch.randn(2, 3, 5, 5)
    torch.testing.assert_close(scaled_masked_softmax(s4, mask, 2.0), torch.softmax((s4 * 2.0).masked_fill(mask.bool(), -10000.0), -1))
    t, freqs = torch.randn(6, 2, 3, 8), torch.randn(6, 1, 1, 8)
    cos, sin = freqs.cos(), freqs.sin()
    rot = torch.cat((-t[..., 4:], t[..., :4]), -1)
    torch.testing.assert_close(fused_apply_rotary_pos_emb(t, freqs), t * cos + rot * sin, atol=1e-5, rtol=1e-5)


def test_fmha_varlen_matches_per_sequence_attention():
    from types import SimpleNamespace

    from apex_b200.contrib.fmha import FMHA
    torch.manual_seed(0)
    h, d = 4, 16
    lens = [5, 9, 1, 12]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    qkv = torch.randn(sum(lens), 3 * h * d, requires_grad=True)
    m = FMHA(SimpleNamespace(attention_probs_dropout_prob=0.0, num_attention_heads=h, hidden_size=h * d))
    out = m(qkv, cu, max(lens), is_training=False)
    assert out.shape == (sum(lens), h * d)
    q3 = qkv.view(-1, 3, h, d)
    for i, n in enumerate(lens):
        s = int(cu[i])
        q, k, v = (q3[s:s + n, j].transpose(0, 1) for j in range(3))   # [h, n, d]
        ref = torch.softmax(q @ k.transpose(1, 2) / d ** 0.5, -1) @ v
        torch.testing.assert_close(out[s:s + n].view(n, h, d).transpose(0, 1), ref, atol=1e-5, rtol=1e-5)
    out.sum().backward()
    assert torch.isfinite(qkv.grad).all()


def test_gds_file_roundtrip(tmp_path):
    from apex_b200.contrib.gpu_direct_storage import GDSFile
    a, b = torch.randn(7, 3), torch.randn(5).bfloat16()
    path = str(tmp_path / "blob.bin")
    with GDSFile(path, "w") as f:
        f.save_data(a)
        f.save_data(b)
    a2, b2 = torch.empty_like(a), torch.empty_like(b)
    with GDSFile(path, "r") as f:
        f.load_data(a2)
        f.load_data(b2)
    assert torch.equal(a, a2) and torch.equal(b, b2)
    # non-contiguous tensors, "rw" mode, reading past the end, a missing file
    c = torch.arange(24.0).view(4, 6).t()
    with GDSFile(path, "rw") as f:
        f.save_data(c)
        f._pos = 0
        c2 = torch.empty(4, 6).t()
        f.load_data(c2)
        assert torch.equal(c, c2)
    with GDSFile(path, "r") as f, pytest.raises(EOFError):
        f.load_data(torch.empty(1 << 20))
    with GDSFile(str(tmp_path / "missing.bin"), "r") as f, pytest.raises(OSError):
        f.load_data(torch.empty(4))


def test_wgrad_accumulation_and_scale_mask_softmax_module_cpu():
    from apex_b200.transformer.functional import FusedScaleMaskSoftmax
    from apex_b200.transformer.functional import fused_weight_gradient as W
    from apex_b200.transformer.functional.fused_softmax import AttnMaskType
    torch.manual_seed(0)
    x, dy = torch.randn(6, 4, 8), torch.randn(6, 4, 5)
    main_grad = torch.ones(5, 8)
    W.wgrad_gemm_accum_fp32(x, dy, main_grad)   # beta = 1 accumulation into the persistent main-grad buffer
    torch.testing.assert_close(main_grad, 1 + dy.reshape(-1, 5).t() @ x.reshape(-1, 8))
    m = FusedScaleMaskSoftmax(False, False, AttnMaskType.padding, True, lambda s, mk: s.masked_fill(mk, -10000.0), True, 0.5)
    s = torch.randn(2, 3, 5, 5)
    mk = torch.zeros(2, 1, 5, 5, dtype=torch.bool)
    mk[..., 4] = True
    torch.testing.assert_close(m(s, mk), torch.softmax((s * 0.5).masked_fill(mk, -10000.0), -1))
    # 16-bit input + fusion enabled -> the fused (here: CPU oracle) causal path, which masks implicitly
    causal = FusedScaleMaskSoftmax(False, True, AttnMaskType.causal, True, None, True, None)
    tri = torch.triu(torch.ones(5, 5, dtype=torch.bool), 1)
    sb = s.bfloat16()
    torch.testing.assert_close(causal(sb, None).float(), torch.softmax(sb.float().masked_fill(tri, float("-inf")), -1), atol=1e-2, rtol=1e-2)


def test_extension_name_modules():
    """`install_as_apex` registers the reference's extension names; their raw entry points follow the reference's signatures."""
    import importlib

    import apex_b200

    apex_b200.install_as_apex()
    ln = importlib.import_module("fused_layer_norm_cuda")
    x, w, b = torch.randn(4, 8), torch.randn(8), torch.randn(8)
    y, mean, invvar = ln.forward_affine(x, (8,), w, b, 1e-5)
    torch.testing.assert_close(y, torch.nn.functional.layer_norm(x, (8,), w, b, 1e-5))
    assert mean.shape == invvar.shape == (4,)
    y, invvar = ln.rms_forward_affine(x, (8,), w, 1e-5)
    torch.testing.assert_close(y, x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * w)
    sm = importlib.import_module("scaled_upper_triang_masked_softmax_cuda")
    a = torch.randn(2, 5, 5)
    p = sm.forward(a, 0.5)
    mask = torch.triu(torch.ones(5, 5, dtype=torch.bool), 1)
    torch.testing.assert_close(p, torch.softmax((a * 0.5).masked_fill(mask, float("-inf")), -1))
    assert sm.backward(torch.randn(2, 5, 5), p, 0.5).shape == a.shape
    xe = importlib.import_module("xentropy_cuda")
    lg, lab = torch.randn(6, 10), torch.randint(0, 10, (6,))
    losses, mlse = xe.forward(lg, lab, 0.1, True)
    torch.testing.assert_close(losses, torch.nn.functional.cross_entropy(lg, lab, label_smoothing=0.1, reduction="none"))
    assert xe.backward(torch.ones(6), lg, mlse, lab, 0.1).shape == lg.shape
    rope = importlib.import_module("fused_rotary_positional_embedding")
    t5 = torch.randn(2, 3, 4, 2, 8)                       # the raw 2-D entry point takes [b, img_h, img_w, heads, d]
    ch, sh, cw, sw = torch.randn(1, 4, 1, 4), torch.randn(1, 4, 1, 4), torch.randn(1, 6, 1, 4), torch.randn(1, 6, 1, 4)
    from apex_b200.transformer.functional import fused_rope as R
    torch.testing.assert_close(rope.forward_2d(t5, ch, sh, cw, sw), R.fused_apply_rotary_pos_emb_2d(t5.view(2, 12, 2, 8), 3, 4, ch, sh, cw, sw).view(2, 3, 4, 2, 8))
    assert rope.backward_2d(t5, ch, sh, cw, sw).shape == t5.shape
    assert rope.forward_thd(torch.randn(12, 2, 8), torch.tensor([0, 5, 12], dtype=torch.int32), torch.randn(8, 1, 1, 8)).shape == (12, 2, 8)
    for name in ("amp_C", "syncbn", "apex_C", "fused_weight_gradient_mlp_cuda", "scaled_masked_softmax_cuda", "scaled_softmax_cuda",
                 "generic_scaled_masked_softmax_cuda", "fused_rotary_positional_embedding"):
        assert importlib.import_module(name) is not None


def test_extension_name_modules_dense_and_distopt():
    """fused_dense_cuda / mlp_cuda / fast_layer_norm / distributed_adam_cuda / distributed_lamb_cuda raw entry points on the CPU paths."""
    import importlib

    import apex_b200

    apex_b200.install_as_apex()
    torch.manual_seed(0)
    fd, mlp_cuda = importlib.import_module("fused_dense_cuda"), importlib.import_module("mlp_cuda")
    x, w1, b1, w2, b2 = torch.randn(8, 16), torch.randn(32, 16), torch.randn(32), torch.randn(4, 32), torch.randn(4)
    leaves = [t.clone().requires_grad_() for t in (x, w1, b1, w2, b2)]
    ref = F.linear(F.gelu(F.linear(leaves[0], leaves[1], leaves[2])), leaves[3], leaves[4])
    dy = torch.randn_like(ref)
    ref.backward(dy)
    o1, o2, gelu_in = fd.linear_gelu_linear_forward(x, w1, b1, w2, b2)
    torch.testing.assert_close(o2, ref.detach())
    for got, leaf in zip(fd.linear_gelu_linear_backward(x, gelu_in, o1, w1, w2, dy), leaves):
        torch.testing.assert_close(got, leaf.grad, rtol=1e-4, atol=1e-4)
    for t in leaves:
        t.grad = None
    ref = torch.sigmoid(F.linear(torch.sigmoid(F.linear(leaves[0], leaves[1], leaves[2])), leaves[3], leaves[4]))
    ref.backward(dy)
    outs = mlp_cuda.forward(1, 2, [x, w1, w2, b1, b2])
    torch.testing.assert_close(outs[0], ref.detach())
    grads = mlp_cuda.backward(1, 2, dy, outs, [x, w1, w2, b1, b2])
    for got, leaf in zip(grads, [leaves[0], leaves[1], leaves[3], leaves[2], leaves[4]]):
        torch.testing.assert_close(got, leaf.grad, rtol=1e-4, atol=1e-4)
    fln = importlib.import_module("fast_layer_norm")
    g, b = torch.randn(16), torch.randn(16)
    z, mu, rs = fln.ln_fwd(x, g, b, 1e-5)
    torch.testing.assert_close(z, F.layer_norm(x, (16,), g, b, 1e-5))

    # distributed_adam_cuda: plain and bf16 + int16-remainder variants against torch.optim.AdamW on the fp32 master
    da = importlib.import_module("distributed_adam_cuda")
    p0 = torch.randn(1000)
    grads = [torch.randn(1000) for _ in range(3)]
    pr = p0.clone().requires_grad_()
    opt = torch.optim.AdamW([pr], lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.1)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    bits = p0.view(torch.int32)
    lo = ((bits & 0xFFFF) ^ 0x8000) - 0x8000
    hi16 = ((bits - lo) >> 16).to(torch.int16)
    p_bf16, rem = hi16.view(torch.bfloat16).clone(), lo.to(torch.int16)
    m2, v2 = torch.zeros(1000), torch.zeros(1000)
    noop, one = torch.zeros(1, dtype=torch.int32), torch.ones(1)
    for step, gr in enumerate(grads, 1):
        pr.grad = gr.clone()
        opt.step()
        da.multi_tensor_fused_adam(65536, noop, [[p], [m], [v], [gr.clone()], [p]], one, 1e-2, 0.9, 0.99, 1e-8, step, 1, 1, 0.1)
        da.multi_tensor_fused_adam_with_param_remainders(65536, noop, [[p_bf16], [rem], [m2], [v2], [gr.clone()], [p_bf16]], one, 1e-2, 0.9, 0.99,
                                                         1e-8, step, 1, 1, 0.1)
    torch.testing.assert_close(p, pr.detach(), rtol=1e-5, atol=1e-6)
    master = ((p_bf16.view(torch.int16).to(torch.int32) << 16) + rem.to(torch.int32)).view(torch.float32)
    torch.testing.assert_close(master, pr.detach(), rtol=1e-5, atol=1e-6)          # hi + remainder IS the fp32 master
    torch.testing.assert_close(p_bf16.float(), pr.detach(), rtol=1e-2, atol=1e-2)  # and hi alone is its bf16 rounding

    # distributed_lamb_cuda: update term + weight update == the multi-tensor LAMB oracle with one hyper-parameter set
    from apex_b200.ops import reference as oracle

    dl = importlib.import_module("distributed_lamb_cuda")
    ps = [torch.randn(50), torch.randn(7, 9)]
    gs = [torch.randn_like(t) for t in ps]
    n = len(ps)
    ms, vs, us = ([torch.zeros_like(t) for t in ps] for _ in range(3))
    p_ref, g_ref, m_ref, v_ref = ([t.clone() for t in l] for l in (ps, gs, ms, vs))
    gnorm = torch.sqrt(sum((g * g).sum() for g in gs)).reshape(1)
    full = lambda val, dt=torch.float32: torch.full((n,), val, dtype=dt)  # noqa: E731
    dl.multi_tensor_lamb_compute_update_term(65536, noop, [gs, ps, ms, vs, us], full(0.9), full(0.999), full(0.1), full(1, torch.int32),
                                             torch.tensor([1], dtype=torch.int32), full(1e-6), 1, full(0.01), one, gnorm, 1.0)
    pn = torch.stack([t.norm() for t in ps])
    un = torch.stack([t.norm() for t in us])
    copies = [torch.zeros_like(t, dtype=torch.bfloat16) for t in ps]
    dl.multi_tensor_lamb_update_weights(65536, noop, [us, ps, copies], pn, un, torch.arange(n), torch.tensor([1e-2]), full(0.01), gnorm, False)
    oracle.multi_tensor_lamb([g_ref, p_ref, m_ref, v_ref], 1e-2, 0.9, 0.999, 1e-6, 1, 1, 0.01, 1, 1, gnorm, 1.0, False)
    for a, b_ in zip(ps, p_ref):
        torch.testing.assert_close(a, b_, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(copies[0].float(), ps[0], rtol=1e-2, atol=1e-2)


def test_attention_modules_route_through_the_kernel_entry_point(monkeypatch):
    """Default plumbing of contrib.fmha / contrib.multihead_attn into the tcgen05 attention kernels, with the (GPU-only) kernel call replaced
    by an SDPA stand-in that honours the same [rows, heads, d] / cu_seqlens / key_bias / bias_div contract: layouts, scaling, masks and the
    (time-major rows, batch x heads as heads) addressing must agree with the generic path."""
    from apex_b200.contrib.fmha import fmha as fm
    from apex_b200.contrib.fmha import kernels as K
    from apex_b200.contrib.multihead_attn import SelfMultiheadAttn

    calls = []

    def stand_in(q, k, v, cu_q, cu_k, max_q, max_k, batch, causal, scale, key_bias=None, dropout_p=0.0, bias_div=0):
        calls.append((tuple(q.shape), batch, causal, key_bias is not None, bias_div))
        rows, h, d = q.shape
        scale = d ** -0.5 if scale is None else scale
        if cu_q is None:
            bounds = [(i * (rows // batch), (i + 1) * (rows // batch), i * (k.shape[0] // batch), (i + 1) * (k.shape[0] // batch)) for i in range(batch)]
        else:
            bounds = [(int(cu_q[i]), int(cu_q[i + 1]), int(cu_k[i]), int(cu_k[i + 1])) for i in range(cu_q.numel() - 1)]
        out = torch.empty(rows, h, d, dtype=q.dtype)
        for bi, (a, b_, c, e) in enumerate(bounds):
            qq, kk, vv = q[a:b_].transpose(0, 1), k[c:e].transpose(0, 1), v[c:e].transpose(0, 1)
            mask = None
            if key_bias is not None:   # [h, 1, tk]: row head // bias_div (or the batch index)
                rowsel = torch.arange(h) // bias_div if bias_div else torch.full((h,), bi)
                mask = key_bias[rowsel][:, None, :e - c].to(qq.dtype)
                if causal:
                    mask = mask + torch.zeros(b_ - a, e - c).masked_fill_(torch.ones(b_ - a, e - c, dtype=torch.bool).triu(1), float("-inf"))
            out[a:b_] = F.scaled_dot_product_attention(qq, kk, vv, attn_mask=mask, is_causal=causal and mask is None, scale=scale).transpose(0, 1)
        return out

    monkeypatch.setattr(K, "supported", lambda t, d: d in (64, 128))   # the real gate also demands CUDA fp16 / bf16 and the native library
    monkeypatch.setattr(K.FmhaFunc, "apply", staticmethod(stand_in))
    torch.manual_seed(0)
    lens = [5, 17, 9]
    cu = torch.tensor([0, 5, 22, 31], dtype=torch.int32)
    qkv = torch.randn(31, 3, 2, 64)
    got = fm.fmha_varlen(qkv, cu, max(lens), 0.0, False)
    torch.testing.assert_close(got, fm._generic_varlen(qkv, cu, max(lens), 0.0, False), rtol=1e-4, atol=1e-5)
    got_c = fm.fmha_varlen(qkv, cu, max(lens), 0.0, False, causal=True)
    torch.testing.assert_close(got_c, fm._generic_varlen(qkv, cu, max(lens), 0.0, True), rtol=1e-4, atol=1e-5)
    mha = SelfMultiheadAttn(128, 2, dropout=0.0, bias=True, impl="fast").eval()
    x = torch.randn(7, 3, 128)
    kpm = torch.zeros(3, 7, dtype=torch.bool)
    kpm[1, 5:] = True
    causal = torch.ones(7, 7, dtype=torch.bool).triu(1)
    outs = [mha(x, is_training=False)[0], mha(x, key_padding_mask=kpm, is_training=False)[0], mha(x, attn_mask=causal, is_training=False)[0]]
    monkeypatch.setattr(K, "supported", lambda t, d: False)
    refs = [mha(x, is_training=False)[0], mha(x, key_padding_mask=kpm, is_training=False)[0], mha(x, attn_mask=causal, is_training=False)[0]]
    for o, r in zip(outs, refs):
        torch.testing.assert_close(o, r, rtol=1e-4, atol=1e-5)
    assert calls[0] == ((31, 2, 64), None, False, False, 0) and calls[1][2] is True
    assert calls[2] == ((7, 6, 64), 1, False, False, 2) and calls[3] == ((7, 6, 64), 1, False, True, 2) and calls[4] == ((7, 6, 64), 1, True, False, 2)


def test_fmha_dropout_mask_reference_has_the_requested_keep_rate():
    from apex_b200.contrib.fmha import kernels as K

    m = K.dropout_keep_mask(2, 3, 65, 130, 0.25, (1234, 8))
    assert m.shape == (2, 3, 65, 130) and abs(m.float().mean().item() - 0.75) < 0.02
    assert not torch.equal(m, K.dropout_keep_mask(2, 3, 65, 130, 0.25, (1234, 12)))
    assert K.dropout_keep_mask(1, 1, 4, 4, 0.0, (1, 0)).all()


def test_syncbatchnorm_follows_torch_batchnorm_semantics_sweep():
    """Every combination of input rank (2-D .. 5-D), affine, track_running_stats, momentum (0.1 / None = cumulative average) and
    train / eval over three iterations: outputs, input gradients, running statistics and the batch counter match torch's BatchNorm."""
    import itertools

    from apex_b200.parallel import SyncBatchNorm
    torch.manual_seed(0)
    for shape, affine, track, mom, train in itertools.product([(6, 5), (6, 5, 7), (4, 5, 3, 3), (2, 5, 2, 3, 2)], [True, False], [True, False],
                                                              [0.1, None], [True, False]):
        C = shape[1]
        a = SyncBatchNorm(C, affine=affine, track_running_stats=track, momentum=mom)
        bn = nn.BatchNorm1d if len(shape) <= 3 else nn.BatchNorm2d if len(shape) == 4 else nn.BatchNorm3d
        b = bn(C, affine=affine, track_running_stats=track, momentum=mom)
        a.train(train)
        b.train(train)
        for _ in range(3):
            x = torch.randn(shape) * 2 + 1
            xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
            ya, yb = a(xa), b(xb)
            dy = torch.randn_like(ya)
            ya.backward(dy)
            yb.backward(dy)
            torch.testing.assert_close(ya, yb, atol=1e-5, rtol=1e-4)
            torch.testing.assert_close(xa.grad, xb.grad, atol=1e-5, rtol=1e-4)
            if track:
                torch.testing.assert_close(a.running_mean, b.running_mean, atol=1e-5, rtol=1e-4)
                torch.testing.assert_close(a.running_var, b.running_var, atol=1e-5, rtol=1e-4)
                assert int(a.num_batches_tracked) == int(b.num_batches_tracked)


def test_attention_functional_forms_match_the_modules():
    """The reference's XxxAttnFunc.apply argument lists (self / encdec, fast, norm-add) evaluate the same function as the modules."""
    from apex_b200.contrib.multihead_attn import EncdecMultiheadAttn, SelfMultiheadAttn
    import apex_b200
    apex_b200.install_as_apex()          # the reference's one-file-per-function import paths resolve to contrib/multihead_attn/funcs.py
    from apex.contrib.multihead_attn.encdec_multihead_attn_func import encdec_attn_func
    from apex.contrib.multihead_attn.fast_encdec_multihead_attn_func import fast_encdec_attn_func
    from apex.contrib.multihead_attn.fast_encdec_multihead_attn_norm_add_func import fast_encdec_attn_norm_add_func
    from apex.contrib.multihead_attn.fast_self_multihead_attn_func import fast_self_attn_func
    from apex.contrib.multihead_attn.fast_self_multihead_attn_norm_add_func import fast_self_attn_norm_add_func
    from apex.contrib.multihead_attn.mask_softmax_dropout_func import MaskSoftmaxDropout, fast_mask_softmax_dropout_func
    from apex.contrib.multihead_attn.self_multihead_attn import SelfMultiheadAttn as _SameClass, jit_dropout_add
    from apex.contrib.multihead_attn.self_multihead_attn_func import self_attn_func
    assert _SameClass is SelfMultiheadAttn
    torch.manual_seed(0)
    E, H, T, B = 32, 4, 6, 2
    x, mem = torch.randn(T, B, E), torch.randn(5, B, E)
    kpm = torch.zeros(B, T, dtype=torch.bool)
    kpm[:, -2:] = True
    tm = torch.triu(torch.ones(T, T, dtype=torch.bool), 1)
    m = SelfMultiheadAttn(E, H, bias=True).eval()
    with torch.no_grad():
        m.in_proj_bias.normal_()
        m.out_proj_bias.normal_()
    w = (m.in_proj_weight, m.out_proj_weight, m.in_proj_bias, m.out_proj_bias)
    torch.testing.assert_close(self_attn_func(False, False, H, m.scaling, x, *w, None, False, 0.0), m(x, is_training=False)[0])
    torch.testing.assert_close(fast_self_attn_func(False, False, H, x, *w, kpm, False, 0.0), m(x, key_padding_mask=kpm, is_training=False)[0])
    torch.testing.assert_close(self_attn_func(True, False, H, m.scaling, x, *w, tm, False, 0.0), m(x, attn_mask=tm, is_training=False)[0])
    n = SelfMultiheadAttn(E, H, include_norm_add=True).eval()
    with torch.no_grad():
        n.lyr_nrm_gamma_weights.normal_()
        n.lyr_nrm_beta_weights.normal_()
    got = fast_self_attn_norm_add_func(False, False, H, x, n.lyr_nrm_gamma_weights, n.lyr_nrm_beta_weights, n.in_proj_weight, n.out_proj_weight, None, 0.0)
    torch.testing.assert_close(got, n(x, is_training=False)[0])
    e = EncdecMultiheadAttn(E, H, bias=True).eval()
    got = encdec_attn_func(False, False, H, e.scaling, x, mem, e.in_proj_weight_q, e.in_proj_weight_kv, e.out_proj_weight, e.in_proj_bias_q,
                           e.in_proj_bias_kv, e.out_proj_bias, None, 0.0)
    torch.testing.assert_close(got, e(x, mem, is_training=False)[0])
    e2 = EncdecMultiheadAttn(E, H, include_norm_add=True).eval()
    got = fast_encdec_attn_norm_add_func(False, False, H, x, mem, e2.lyr_nrm_gamma_weights, e2.lyr_nrm_beta_weights, e2.in_proj_weight_q,
                                         e2.in_proj_weight_kv, e2.out_proj_weight, None, 0.0)
    torch.testing.assert_close(got, e2(x, mem, is_training=False)[0])
    e3 = EncdecMultiheadAttn(E, H).eval()
    torch.testing.assert_close(fast_encdec_attn_func(False, False, H, x, mem, e3.in_proj_weight_q, e3.in_proj_weight_kv, e3.out_proj_weight, None, 0.0),
                               e3(x, mem, is_training=False)[0])
    s = torch.randn(B * H, T, T)
    torch.testing.assert_close(MaskSoftmaxDropout.apply(False, H, s, None, False, 0.0), fast_mask_softmax_dropout_func(False, H, s, None, False, 0.0))
    torch.testing.assert_close(jit_dropout_add(x, x, 0.5, False), 2 * x)


def test_bottleneck_scale_bias_callable():
    """get_scale_bias_callable(): forward() reads persistent folded (scale, bias) tensors that only change when the callable runs."""
    from apex_b200.contrib.bottleneck import Bottleneck
    torch.manual_seed(0)
    for explicit in (False, True):
        m = Bottleneck(16, 8, 32, stride=2, explicit_nhwc=explicit)
        for bn in (m.bn1, m.bn2, m.bn3, m.downsample[1]):
            bn.weight.normal_()
            bn.bias.normal_()
            bn.running_mean.normal_()
            bn.running_var.uniform_(0.5, 2)
        x = torch.randn(2, 16, 8, 8)
        x = x.permute(0, 2, 3, 1).contiguous() if explicit else x
        y0 = m(x)
        refresh = m.get_scale_bias_callable()
        refresh()
        torch.testing.assert_close(m(x), y0)
        m.bn1.weight.mul_(2)
        torch.testing.assert_close(m(x), y0)          # stale on purpose until refreshed
        refresh()
        assert not torch.allclose(m(x), y0)


def test_bottleneck_function_matches_the_module():
    """BottleneckFunction.apply(nhwc, stride_1x1, scale, bias, x, *conv) (reference bottleneck.py:80-132) == Bottleneck.forward, values and
    gradients, with and without the downsample branch, NCHW and explicit NHWC (weights [K, R, S, C] there)."""
    from apex_b200.contrib.bottleneck import Bottleneck
    from apex_b200.contrib.bottleneck.bottleneck import BottleneckFunction, SpatialBottleneckFunction, bottleneck_function
    torch.manual_seed(0)
    for cin, cout, stride in ((16, 32, 2), (32, 32, 1)):
        m = Bottleneck(cin, 8, cout, stride=stride)
        norms = [m.bn1, m.bn2, m.bn3] + ([m.downsample[1]] if m.downsample is not None else [])
        for bn in norms:
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2)
        convs = [m.conv1.weight, m.conv2.weight, m.conv3.weight] + ([m.downsample[0].weight] if m.downsample is not None else [])
        scale, bias = zip(*(bn.get_scale_bias() for bn in norms))
        x = torch.randn(2, cin, 8, 8, requires_grad=True)
        want = m(x)
        got = bottleneck_function(False, (stride, stride), list(scale), list(bias), x, *convs)
        torch.testing.assert_close(got, want)
        gw = torch.autograd.grad(want.sum(), [x, m.conv2.weight])
        gg = torch.autograd.grad(got.sum(), [x, m.conv2.weight])
        for a, b in zip(gg, gw):
            torch.testing.assert_close(a, b)
        nhwc = BottleneckFunction.apply(True, (stride, stride), list(scale), list(bias), x.detach().permute(0, 2, 3, 1),
                                        *[w.detach().permute(0, 2, 3, 1) for w in convs])
        torch.testing.assert_close(nhwc, want.detach().permute(0, 2, 3, 1))
        one = SpatialBottleneckFunction.apply(1, 0, None, None, 1, False, False, (stride, stride), list(scale), list(bias), None, None, x, *convs)
        torch.testing.assert_close(one, want)


def test_norm_custom_ops_numerics_opcheck_and_fullgraph_compile(monkeypatch):
    """apex_b200::norm_fwd / norm_bwd (normalization/custom_ops.py): values and gradients against torch, torch.library.opcheck (schema, fake
    implementation, autograd registration, AOT dispatch), a fullgraph torch.compile, and the modules' compile-time route (forced on for CPU
    tensors here; on a GPU it is taken whenever a CUDA graph is being compiled)."""
    from apex_b200.normalization import FusedLayerNorm, FusedRMSNorm, custom_ops as C
    import importlib

    M = importlib.import_module("apex_b200.normalization.fused_layer_norm")   # the package re-exports a function of the same name
    torch.manual_seed(0)
    x, w, b = torch.randn(6, 4, 8, requires_grad=True), torch.randn(8, requires_grad=True), torch.randn(8, requires_grad=True)
    g = torch.randn(6, 4, 8)
    cases = [(C.norm(x, w, b, (8,), 1e-5), F.layer_norm(x, (8,), w, b, 1e-5), (x, w, b)),
             (C.norm(x, w, None, (8,), 1e-5, rms=True), x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * w, (x, w)),
             (C.norm(x, None, None, (4, 8), 1e-5), F.layer_norm(x, (4, 8), None, None, 1e-5), (x,))]
    for got, want, leaves in cases:
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
        for a, r in zip(torch.autograd.grad(got, leaves, g), torch.autograd.grad(want, leaves, g)):
            torch.testing.assert_close(a, r, rtol=1e-4, atol=1e-5)
    y, mean, invvar = C.reference_fwd(x.detach(), (8,), w.detach(), b.detach(), 1e-5, False, torch.float32)
    plain = C.reference_bwd(g, x.detach(), mean, invvar, (8,), w.detach(), b.detach(), False, False, torch.float32)
    from_output = C.reference_bwd(g, y, None, invvar, (8,), w.detach(), b.detach(), False, True, torch.float32)   # memory-efficient: x-hat from y
    for a, r in zip(plain, from_output):
        torch.testing.assert_close(a, r, rtol=1e-4, atol=1e-5)
    for args in [(x.detach().requires_grad_(), w.detach().requires_grad_(), b.detach().requires_grad_(), [8], 1e-5, False, False),
                 (x.detach().requires_grad_(), w.detach().requires_grad_(), None, [8], 1e-5, True, False)]:
        assert set(torch.library.opcheck(C.norm_fwd_op, args).values()) == {"SUCCESS"}
    monkeypatch.setattr(M, "_compiled_cuda", lambda t: torch.compiler.is_compiling())
    net = torch.nn.Sequential(torch.nn.Linear(8, 8), FusedLayerNorm(8), torch.nn.Tanh(), FusedRMSNorm(8))
    compiled = torch.compile(net, backend="aot_eager", fullgraph=True)   # fullgraph: the custom ops do not break the graph
    xin = torch.randn(5, 8, requires_grad=True)
    out, want = compiled(xin), net(xin)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(torch.autograd.grad(out.sum(), xin)[0], torch.autograd.grad(want.sum(), xin)[0], rtol=1e-4, atol=1e-5)


def test_group_norm_custom_ops(monkeypatch):
    """apex_b200::group_norm_nhwc_fprop / _bprop: values and gradients (with and without SiLU) against torch, opcheck, and the module's
    compile-time route under a fullgraph torch.compile."""
    import importlib

    M = importlib.import_module("apex_b200.contrib.group_norm.group_norm")
    torch.manual_seed(0)
    for act in ("silu", ""):
        x = torch.randn(2, 8, 3, 3).contiguous(memory_format=torch.channels_last).requires_grad_()
        w, b = torch.randn(8, requires_grad=True), torch.randn(8, requires_grad=True)
        y = M.group_norm_nhwc_fprop_op(x, 4, w, b, 1e-5, act)[0]
        ref = F.group_norm(x, 4, w, b, 1e-5)
        ref = F.silu(ref) if act else ref
        g = torch.randn_like(y)
        torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-5)
        for a, r in zip(torch.autograd.grad(y, (x, w, b), g), torch.autograd.grad(ref, (x, w, b), g)):
            torch.testing.assert_close(a, r, rtol=1e-4, atol=1e-5)
    args = (x.detach().requires_grad_(), 4, w.detach().requires_grad_(), b.detach().requires_grad_(), 1e-5, "silu")
    assert set(torch.library.opcheck(M.group_norm_nhwc_fprop_op, args).values()) == {"SUCCESS"}
    monkeypatch.setattr(M, "_compiled_cuda", lambda t: torch.compiler.is_compiling())
    net = torch.nn.Sequential(torch.nn.Conv2d(8, 8, 1), M.GroupNorm(4, 8, act="silu")).to(memory_format=torch.channels_last)
    compiled = torch.compile(net, backend="aot_eager", fullgraph=True)
    xin = torch.randn(2, 8, 3, 3).contiguous(memory_format=torch.channels_last).requires_grad_()
    out, want = compiled(xin), net(xin)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(torch.autograd.grad(out.sum(), xin)[0], torch.autograd.grad(want.sum(), xin)[0], rtol=1e-4, atol=1e-5)


def test_state_dict_keys_match_the_reference_constructors():
    """Construct the same modules from the reference package (python only, extensions stubbed) and from this one: parameter / buffer names and
    shapes must agree wherever the reference constructor runs without a GPU, so model checkpoints move in both directions."""
    import json
    import os
    import subprocess
    import sys

    if not os.path.isdir("/root/reference/apex"):
        pytest.skip("reference tree not available")
    helper = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_state_dict_keys.py")
    out = {}
    for which in ("ref", "ours"):
        r = subprocess.run([sys.executable, helper, which], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[which] = json.loads(r.stdout.strip().splitlines()[-1])
    compared = 0
    for case, ref_keys in out["ref"].items():
        if isinstance(ref_keys, str):      # the reference constructor needs CUDA / an initialised process group
            continue
        assert out["ours"][case] == ref_keys, case
        compared += 1
    assert compared >= 15


def test_extension_name_modules_losses_and_layer_norm_backward():
    """focal_loss_cuda / fused_index_mul_2d / fused_layer_norm_cuda.backward* / fast_layer_norm.ln_bwd shims on the PyTorch paths."""
    import importlib

    import apex_b200
    from apex_b200.contrib.focal_loss.focal_loss import _ref as focal_ref

    apex_b200.install_as_apex()
    torch.manual_seed(0)
    fl = importlib.import_module("focal_loss_cuda")
    co, tg, npos = torch.randn(4, 30, 10), torch.randint(-2, 10, (4, 30)), torch.tensor([11.0])
    loss, pgrad = fl.forward(co, tg, npos, 8, 0.25, 2.0, 0.0)
    cr = co.clone().requires_grad_()
    ref = focal_ref(cr, tg, npos, 8, 0.25, 2.0, 0.0)
    ref.backward()
    torch.testing.assert_close(loss, ref.detach())
    torch.testing.assert_close(fl.backward(torch.tensor(2.0), pgrad, npos), 2.0 * cr.grad)
    im = importlib.import_module("fused_index_mul_2d")
    in1, in2, idx = torch.randn(7, 5), torch.randn(20, 5), torch.randint(0, 7, (20,))
    out = torch.empty_like(in2)
    im.float_forward(out, in1, in2, idx)
    torch.testing.assert_close(out, in1[idx] * in2)
    g1, g2, go = torch.zeros_like(in1), torch.empty_like(in2), torch.randn_like(in2)
    im.float_backward(g1, g2, go, in1, in2, idx)
    torch.testing.assert_close(g2, go * in1[idx])
    torch.testing.assert_close(g1, torch.zeros_like(in1).index_add_(0, idx, go * in2))
    ln = importlib.import_module("fused_layer_norm_cuda")
    x, w, b = torch.randn(6, 16), torch.randn(16), torch.randn(16)
    leaves = [t.clone().requires_grad_() for t in (x, w, b)]
    want = F.layer_norm(leaves[0], (16,), leaves[1], leaves[2], 1e-5)
    dy = torch.randn_like(want)
    y, mean, invvar = ln.forward_affine(x, (16,), w, b, 1e-5)
    for got, r in zip(ln.backward_affine(dy, mean, invvar, x, (16,), w, b, 1e-5), torch.autograd.grad(want, leaves, dy)):
        torch.testing.assert_close(got, r, rtol=1e-4, atol=1e-5)
    plain = ln.backward_affine(dy, mean, invvar, x, (16,), w, b, 1e-5)
    for got, r in zip(ln.backward_affine(dy, mean, invvar, y, (16,), w, b, 1e-5, True), plain):
        torch.testing.assert_close(got, r, rtol=1e-4, atol=1e-5)          # memory-efficient form (saved OUTPUT) == plain form (saved input)
    yr, iv = ln.rms_forward_affine(x, (16,), w, 1e-5)
    wr = [t.clone().requires_grad_() for t in (x, w)]
    want_r = wr[0] * torch.rsqrt(wr[0].pow(2).mean(-1, keepdim=True) + 1e-5) * wr[1]
    for got, r in zip(ln.rms_backward_affine(dy, iv, x, (16,), w, 1e-5), torch.autograd.grad(want_r, wr, dy)):
        torch.testing.assert_close(got, r, rtol=1e-4, atol=1e-5)
    fast = importlib.import_module("fast_layer_norm")
    z, mu, rs = fast.ln_fwd(x, w, b, 1e-5)
    dx, dg, db, _, _ = fast.ln_bwd(dy, x, mu, rs, w)
    for got, r in zip((dx, dg, db), torch.autograd.grad(F.layer_norm(leaves[0], (16,), leaves[1], leaves[2], 1e-5), leaves, dy)):
        torch.testing.assert_close(got, r, rtol=1e-4, atol=1e-5)
