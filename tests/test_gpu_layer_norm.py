"""FusedLayerNorm / FusedRMSNorm kernels vs fp32 PyTorch (reference tests/L0/run_fused_layer_norm)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [((16, 1024), 1024), ((3, 7, 768), 768), ((64, 4096), 4096), ((5, 8192), 8192), ((4, 16384), 16384), ((33, 100), 100),
          ((17, 65), 65), ((2, 3, 24), 24), ((9, 32768), 32768), ((301, 12288), 12288), ((333, 16384), 16384),
          ((40, 2560), 2560), ((37, 5120), 5120), ((21, 6144), 6144), ((19, 7168), 7168), ((11, 10240), 10240), ((9, 14336), 14336)]   # the last two: more rows than
# resident CTAs / clusters (persistent row loops, the 2-CTA row split of the backward); then LLM hidden sizes that are not a power-of-two multiple of the
# thread count (the last vectors of some threads are predicated off)
TOL = {torch.float32: (1e-5, 1e-4), torch.float16: (2e-3, 2e-2), torch.bfloat16: (2e-2, 1e-1)}


def _ref(x, w, b, n2, eps, rms):
    xf = x.float()
    if rms:
        y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
        return y * w.float() if w is not None else y
    return F.layer_norm(xf, (n2,), None if w is None else w.float(), None if b is None else b.float(), eps)


@pytest.mark.parametrize("shape,n2", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("rms", [False, True])
@pytest.mark.parametrize("memory_efficient", [False, True])
def test_norm_fwd_bwd(cuda_dev, shape, n2, dtype, rms, memory_efficient):
    from apex_b200.normalization import FusedLayerNorm, FusedRMSNorm
    torch.manual_seed(0)
    mod = (FusedRMSNorm if rms else FusedLayerNorm)(n2, eps=1e-5, memory_efficient=memory_efficient).to(cuda_dev, dtype)
    with torch.no_grad():
        w = torch.randn(n2) * 0.5 + 1.0
        if memory_efficient:  # x-hat is rebuilt as (y-b)/w: keep |w| away from 0 or 16-bit rounding is amplified without bound
            w = w.abs() + 0.5
        mod.weight.copy_(w)
        if not rms:
            mod.bias.copy_(torch.randn(n2) * 0.1)
    x = torch.randn(shape, device=cuda_dev, dtype=dtype, requires_grad=True)
    dy = torch.randn(shape, device=cuda_dev, dtype=dtype)
    y = mod(x)
    y.backward(dy)
    xr = x.detach().clone().float().requires_grad_(True)
    wr = mod.weight.detach().clone().float().requires_grad_(True)
    br = None if rms else mod.bias.detach().clone().float().requires_grad_(True)
    yr = _ref(xr, wr, br, n2, 1e-5, rms)
    yr.backward(dy.float())
    ftol, btol = TOL[dtype]
    torch.testing.assert_close(y.float(), yr, atol=ftol, rtol=ftol)
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=btol, rtol=btol)
    rows = x.numel() // n2
    gscale = max(1.0, rows ** 0.5)
    torch.testing.assert_close(mod.weight.grad.float(), wr.grad, atol=btol * gscale, rtol=btol)
    if not rms:
        torch.testing.assert_close(mod.bias.grad.float(), br.grad, atol=btol * gscale, rtol=btol)


def test_no_affine_and_mixed_dtype_and_strided_input(cuda_dev):
    from apex_b200.normalization import FusedLayerNorm, MixedFusedLayerNorm, MixedFusedRMSNorm
    x = torch.randn(6, 10, 12, 256, device=cuda_dev)[::3, ::5, ::3]  # non-contiguous
    m = FusedLayerNorm(256, elementwise_affine=False).to(cuda_dev)
    torch.testing.assert_close(m(x), F.layer_norm(x, (256,)), atol=1e-5, rtol=1e-4)
    xb = torch.randn(32, 1024, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    for M in (MixedFusedLayerNorm, MixedFusedRMSNorm):
        mm = M(1024).to(cuda_dev)  # fp32 params, bf16 activations -> fp32 output
        y = mm(xb)
        assert y.dtype == torch.float32
        y.sum().backward()
        assert xb.grad.dtype == torch.bfloat16 and mm.weight.grad.dtype == torch.float32


def test_autocast(cuda_dev):
    from apex_b200.normalization import FusedLayerNorm
    m = FusedLayerNorm(512).to(cuda_dev)
    x = torch.randn(8, 512, device=cuda_dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    assert y.dtype == torch.bfloat16
    torch.testing.assert_close(y.float(), F.layer_norm(x, (512,)), atol=3e-2, rtol=3e-2)
