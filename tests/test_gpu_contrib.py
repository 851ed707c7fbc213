"""GPU numerics of the contrib kernels against PyTorch fp32 oracles: transducer joint / loss, focal loss, index_mul_2d,
clip_grad, multihead attention. Mirrors apex/contrib/test/{transducer,focal_loss,index_mul_2d,clip_grad,multihead_attn}."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lens(B, T, U, dev):
    g = torch.Generator().manual_seed(3)
    f_len = torch.randint(T // 2, T + 1, (B,), generator=g)
    y_len = torch.randint(U // 2, U, (B,), generator=g)
    f_len[0], y_len[-1] = T, U - 1
    return f_len.to(dev).int(), y_len.to(dev).int()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("pack", [False, True])
def test_transducer_joint(cuda_dev, dtype, relu, pack):
    from apex_b200.contrib.transducer import TransducerJoint
    from apex_b200.contrib.transducer.transducer import _TorchTransducerJoint
    B, T, U, H = 4, 13, 7, 72
    torch.manual_seed(0)
    f = torch.randn(B, T, H, device=cuda_dev, dtype=dtype, requires_grad=True)
    g = torch.randn(B, U, H, device=cuda_dev, dtype=dtype, requires_grad=True)
    f_len, y_len = _lens(B, T, U, cuda_dev)
    g_len = y_len + 1
    bo = torch.cumsum(f_len.long() * g_len.long(), 0)
    pb = int(bo[-1])
    out = TransducerJoint(pack_output=pack, relu=relu)(f, g, f_len, g_len, batch_offset=bo, packed_batch=pb)
    fr, gr = f.detach().float().requires_grad_(True), g.detach().float().requires_grad_(True)
    ref = _TorchTransducerJoint.forward(_TorchTransducerJoint(pack_output=pack, relu=relu), fr, gr, f_len, g_len, batch_offset=bo, packed_batch=pb)
    tol = 1e-5 if dtype == torch.float32 else 2e-3
    torch.testing.assert_close(out.float(), ref, atol=tol, rtol=tol)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    out.backward(dy.to(dtype))
    torch.testing.assert_close(f.grad.float(), fr.grad, atol=tol * 20, rtol=tol * 20)
    torch.testing.assert_close(g.grad.float(), gr.grad, atol=tol * 20, rtol=tol * 20)


def test_transducer_joint_dropout(cuda_dev):
    from apex_b200.contrib.transducer import TransducerJoint
    B, T, U, H = 2, 9, 5, 64
    f = torch.randn(B, T, H, device=cuda_dev, requires_grad=True)
    g = torch.randn(B, U, H, device=cuda_dev, requires_grad=True)
    f_len = torch.tensor([9, 6], device=cuda_dev).int()
    g_len = torch.tensor([5, 3], device=cuda_dev).int()
    j = TransducerJoint(dropout=True, dropout_prob=0.25, probe_mask=True)
    out = j(f, g, f_len, g_len)
    mask = j.mask_probe[0]
    ref = (f.detach().unsqueeze(2) + g.detach().unsqueeze(1)) * mask / 0.75
    valid = (torch.arange(T, device=cuda_dev).view(1, T, 1) < f_len.view(B, 1, 1)) & (torch.arange(U, device=cuda_dev).view(1, 1, U) < g_len.view(B, 1, 1))
    torch.testing.assert_close(out, ref * valid.unsqueeze(-1), atol=1e-5, rtol=1e-5)
    keep = mask[valid].float().mean().item()
    assert 0.70 < keep < 0.80
    out.sum().backward()
    torch.testing.assert_close(f.grad, (mask.float() / 0.75 * valid.unsqueeze(-1)).sum(2), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("packed", [False, True])
def test_transducer_loss(cuda_dev, dtype, packed):
    from apex_b200.contrib.transducer import TransducerLoss
    from apex_b200.contrib.transducer.transducer import _TorchTransducerLoss
    B, T, U, V = 5, 17, 9, 37
    torch.manual_seed(1)
    f_len, y_len = _lens(B, T, U, cuda_dev)
    label = torch.randint(1, V, (B, U - 1), device=cuda_dev).int()
    x = torch.randn(B, T, U, V, device=cuda_dev, dtype=dtype)
    xr = x.detach().float().requires_grad_(True)
    ref = _TorchTransducerLoss.forward(_TorchTransducerLoss(), xr, label, f_len, y_len, 0)
    w = torch.rand(B, device=cuda_dev) + 0.5
    (ref * w).sum().backward()
    valid = (torch.arange(T, device=cuda_dev).view(1, T, 1) < f_len.view(B, 1, 1)) & (torch.arange(U, device=cuda_dev).view(1, 1, U) < (y_len + 1).view(B, 1, 1))
    if packed:
        bo = torch.cumsum(f_len.long() * (y_len.long() + 1), 0)
        xin = x[valid].clone().requires_grad_(True)
        loss = TransducerLoss(packed_input=True)(xin, label, f_len, y_len, 0, batch_offset=bo, max_f_len=T)
    else:
        xin = x.clone().requires_grad_(True)
        loss = TransducerLoss()(xin, label, f_len, y_len, 0)
    tol = 1e-4 if dtype == torch.float32 else 5e-2
    torch.testing.assert_close(loss.float(), ref.detach(), atol=tol, rtol=1e-4 if dtype == torch.float32 else 2e-3)
    (loss * w).sum().backward()
    gref = xr.grad * valid.unsqueeze(-1)
    got = xin.grad.float()
    gtol = 1e-4 if dtype == torch.float32 else 5e-3
    torch.testing.assert_close(got, gref[valid] if packed else gref, atol=gtol, rtol=gtol)


def test_focal_loss(cuda_dev):
    from apex_b200.contrib.focal_loss import focal_loss
    from apex_b200.contrib.focal_loss.focal_loss import _ref
    torch.manual_seed(0)
    N, C = 4096, 91
    x = torch.randn(N, C, device=cuda_dev, requires_grad=True)
    tgt = torch.randint(-2, 80, (N,), device=cuda_dev)
    npos = torch.tensor([123.0], device=cuda_dev)
    loss = focal_loss(x, tgt, npos, 80, 0.25, 2.0, 0.1)
    xr = x.detach().clone().requires_grad_(True)
    ref = _ref(xr, tgt, npos, 80, 0.25, 2.0, 0.1)
    torch.testing.assert_close(loss, ref.reshape(loss.shape), atol=1e-3, rtol=1e-4)
    loss.backward()
    ref.backward()
    torch.testing.assert_close(x.grad, xr.grad, atol=1e-6, rtol=1e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_index_mul_2d(cuda_dev, dtype):
    from apex_b200.contrib.index_mul_2d import index_mul_2d
    torch.manual_seed(0)
    in1 = torch.randn(500, 64, device=cuda_dev, dtype=dtype, requires_grad=True)
    in2 = torch.randn(6000, 64, device=cuda_dev, dtype=dtype, requires_grad=True)
    idx = torch.randint(0, 500, (6000,), device=cuda_dev)
    out = index_mul_2d(in1, in2, idx)
    a, b = in1.detach().float().requires_grad_(True), in2.detach().float().requires_grad_(True)
    ref = a[idx] * b
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    torch.testing.assert_close(out.float(), ref, atol=tol, rtol=tol)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    out.backward(dy.to(dtype))
    torch.testing.assert_close(in1.grad.float(), a.grad, atol=tol * 20, rtol=tol * 5)
    torch.testing.assert_close(in2.grad.float(), b.grad, atol=tol * 5, rtol=tol * 5)


def test_clip_grad(cuda_dev):
    from apex_b200.contrib.clip_grad import clip_grad_norm_
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(n, device=cuda_dev, dtype=dt)) for n, dt in [(1000, torch.float32), (37, torch.float16), (70000, torch.float32)]]
    rs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    for p, r in zip(ps, rs):
        p.grad = torch.randn_like(p)
        r.grad = p.grad.clone()
    n1 = clip_grad_norm_(ps, 0.5)
    n2 = torch.nn.utils.clip_grad_norm_(rs, 0.5)
    torch.testing.assert_close(n1.float().reshape(()), n2.float().reshape(()), atol=1e-3, rtol=1e-3)
    for p, r in zip(ps, rs):
        torch.testing.assert_close(p.grad.float(), r.grad.float(), atol=1e-3, rtol=2e-3)


def test_self_multihead_attn(cuda_dev):
    from apex_b200.contrib.multihead_attn import SelfMultiheadAttn
    torch.manual_seed(0)
    E, Hh, S, B = 128, 8, 24, 3
    m = SelfMultiheadAttn(E, Hh, dropout=0.0, bias=True).to(cuda_dev)
    ref = torch.nn.MultiheadAttention(E, Hh, dropout=0.0, bias=True).to(cuda_dev)
    from apex_b200.contrib.multihead_attn.multihead_attn import packed_to_blocked
    with torch.no_grad():   # packed projection rows are [head][q|k|v][head_dim] (the reference's layout); torch wants [q; k; v] blocks
        ref.in_proj_weight.copy_(packed_to_blocked(m.in_proj_weight, Hh))
        ref.in_proj_bias.copy_(packed_to_blocked(m.in_proj_bias, Hh))
        ref.out_proj.weight.copy_(m.out_proj_weight)
        ref.out_proj.bias.copy_(m.out_proj_bias)
    x = torch.randn(S, B, E, device=cuda_dev)
    out, _ = m(x, x, x, is_training=False)
    r, _ = ref(x, x, x, need_weights=False)
    torch.testing.assert_close(out, r, atol=2e-4, rtol=2e-4)


def test_symmetric_pluggable_allocator_pool(cuda_dev):
    """torch.cuda.MemPool over csrc/symm_heap.cpp ab_symm_pool_malloc / _free: allocations come from shareable VMM blocks."""
    import ctypes
    from apex_b200 import _lib
    from apex_b200.contrib.nccl_allocator import nccl_allocator as A
    blocks = _lib.raw_fn("ab_symm_pool_blocks", ctypes.c_int, [])
    n0 = blocks()
    with A.symmetric_mem():
        a = torch.randn(1 << 20, device=cuda_dev)
        b = torch.randn(1 << 20, device=cuda_dev)
    assert blocks() > n0
    torch.testing.assert_close((a + b).sum(), a.sum() + b.sum(), rtol=1e-4, atol=1e-2)
    base, nbytes, fd = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_int(-1)
    _lib.declare("ab_symm_pool_export", "p p p p")
    _lib.fn("ab_symm_pool_export")(a.data_ptr(), ctypes.addressof(base), ctypes.addressof(nbytes), ctypes.addressof(fd))
    assert base.value <= a.data_ptr() < base.value + nbytes.value and fd.value >= 0
    import os
    os.close(fd.value)


def test_symmetric_pool_peer_map_two_gpus(cuda_dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.symmetric_pool_peer_map, 2, "cuda", backend="nccl")


def test_nccl_p2p_native_communicator(cuda_dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.nccl_p2p_native_exchange, 2, "cuda", backend="nccl")


def test_spatial_bottleneck_two_gpus(cuda_dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.spatial_bottleneck_matches_full, 2, "cuda", backend="nccl")


@pytest.mark.parametrize("world", [2, 4])
def test_peer_halo_exchange(cuda_dev, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.peer_halo_exchange_matches_allgather, world, "cuda", backend="nccl")


# ---------------------------------------------------------------- fused conv tails (csrc/conv_epilogue.cu)
@pytest.mark.parametrize("dtype,C_out", [(torch.bfloat16, 64), (torch.float16, 24), (torch.float32, 32), (torch.float16, 13)])
@pytest.mark.parametrize("variant", ["bias_relu", "bias", "mask", "scale_bias_add_relu"])
def test_fused_conv_epilogue_forward_backward(cuda_dev, dtype, C_out, variant):
    import torch.nn.functional as F
    from apex_b200.contrib.conv_bias_relu.conv_bias_relu import fused_conv_epilogue
    torch.manual_seed(0)
    x = torch.randn(3, 16, 14, 10, device=cuda_dev).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_()
    w = (torch.randn(C_out, 16, 3, 3, device=cuda_dev) * 0.1).to(dtype).requires_grad_()
    b = torch.randn(1, C_out, 1, 1, device=cuda_dev).to(dtype).requires_grad_()
    sc = (torch.rand(C_out, device=cuda_dev) + 0.5).requires_grad_() if variant == "scale_bias_add_relu" else None
    z = torch.randn(3, C_out, 14, 10, device=cuda_dev).to(dtype).requires_grad_() if variant == "scale_bias_add_relu" else None
    mask = (torch.rand(3, C_out, 14, 10, device=cuda_dev) > 0.3) if variant == "mask" else None
    relu = variant != "bias"
    out = fused_conv_epilogue(x, w, bias=b, scale=sc, z=z, mask=mask, stride=1, padding=1, relu=relu)
    # fp32 reference of the same op
    xf, wf, bf = (t.detach().float().requires_grad_() for t in (x, w, b))
    scf = sc.detach().clone().requires_grad_() if sc is not None else None
    zf = z.detach().float().requires_grad_() if z is not None else None
    ref = F.conv2d(xf, wf, None, 1, 1)
    if scf is not None:
        ref = ref * scf.view(1, -1, 1, 1)
    ref = ref + bf
    if zf is not None:
        ref = ref + zf
    if mask is not None:
        ref = ref * mask
    if relu:
        # the ReLU decision is taken from the kernel's own (16-bit) output so that elements within rounding distance of zero do not flip
        # between the two computations; everything else of the reference is fp32
        fwd_ref = F.relu(ref)
        ref = ref * (out.detach().float() > 0)
    else:
        fwd_ref = ref
    tol = 2e-5 if dtype == torch.float32 else 3e-2
    torch.testing.assert_close(out.float(), fwd_ref, atol=tol * 4, rtol=tol)
    g = torch.randn_like(out)
    ins = [t for t in (x, w, b, sc, z) if t is not None]
    refs = [t for t in (xf, wf, bf, scf, zf) if t is not None]
    got = torch.autograd.grad(out, ins, g)
    want = torch.autograd.grad(ref, refs, g.float())
    for a_, r_, name in zip(got, want, ["x", "w", "b", "scale", "z"][:len(got)] if sc is not None else ["x", "w", "b"]):
        scale_ = max(1.0, r_.abs().max().item())
        assert (a_.float() - r_).abs().max().item() <= (5e-4 if dtype == torch.float32 else 6e-2) * scale_, name


def test_bottleneck_block_matches_eager_composition(cuda_dev):
    import torch.nn.functional as F
    from apex_b200.contrib.bottleneck import Bottleneck
    torch.manual_seed(0)
    blk = Bottleneck(32, 16, 64, stride=2).to(cuda_dev).half().to(memory_format=torch.channels_last)
    for bn in (blk.bn1, blk.bn2, blk.bn3, blk.downsample[1]):
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 1.5)  # noqa: E702
    x = torch.randn(4, 32, 16, 16, device=cuda_dev).half().contiguous(memory_format=torch.channels_last).requires_grad_()
    out = blk(x)

    def eager(xx):
        f = lambda bn: tuple(t.float() for t in bn.get_scale_bias())  # noqa: E731
        (s1, b1), (s2, b2), (s3, b3), (s4, b4) = f(blk.bn1), f(blk.bn2), f(blk.bn3), f(blk.downsample[1])
        w = lambda c: c.weight.float()  # noqa: E731
        o = F.relu(F.conv2d(xx, w(blk.conv1), None, 2) * s1 + b1)
        o = F.relu(F.conv2d(o, w(blk.conv2), None, 1, 1) * s2 + b2)
        o = F.conv2d(o, w(blk.conv3)) * s3 + b3
        return F.relu(o + F.conv2d(xx, w(blk.downsample[0]), None, 2) * s4 + b4)

    xf = x.detach().float().requires_grad_()
    ref = eager(xf)
    torch.testing.assert_close(out.float(), ref, atol=5e-2, rtol=5e-2)
    g = torch.randn_like(out)
    (gx,) = torch.autograd.grad(out, x, g)
    (rx,) = torch.autograd.grad(ref, xf, g.float())
    assert (gx.float() - rx).abs().max().item() <= 5e-2 * max(1.0, rx.abs().max().item())
