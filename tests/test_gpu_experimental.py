"""Experimental kernels (csrc/experimental): only run when the library was built with APEX_B200_EXPERIMENTAL=1 — skipped otherwise."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fmha():
    from apex_b200.contrib.fmha import experimental as X
    if not X.available():
        pytest.skip("experimental kernels are not in this build (APEX_B200_EXPERIMENTAL=1 python -m apex_b200._build)")
    return X


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("seq", [128, 200, 517])
def test_fmha_fwd_fixed_length(cuda_dev, d, causal, seq):
    X = _fmha()
    torch.manual_seed(0)
    b, h = 3, 4
    qkv = torch.randn(b * seq, 3, h, d, device=cuda_dev, dtype=torch.bfloat16)
    out, lse = X.fmha_fwd(qkv[:, 0], qkv[:, 1], qkv[:, 2], batch=b, causal=causal, return_lse=True)
    q, k, v = (qkv[:, i].view(b, seq, h, d).transpose(1, 2).float() for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal)
    torch.testing.assert_close(out.view(b, seq, h, d).transpose(1, 2).float(), ref, atol=2e-2, rtol=2e-2)


def test_fmha_fwd_varlen(cuda_dev):
    X = _fmha()
    torch.manual_seed(0)
    h, d = 4, 64
    lens = [5, 130, 1, 300]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=cuda_dev)
    qkv = torch.randn(sum(lens), 3, h, d, device=cuda_dev, dtype=torch.float16)
    out = X.fmha_fwd(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu_seqlens_q=cu, max_seqlen_q=max(lens))
    for i, n in enumerate(lens):
        s = int(cu[i])
        q, k, v = (qkv[s:s + n, j].transpose(0, 1).float() for j in range(3))
        ref = torch.softmax(q @ k.transpose(1, 2) / d ** 0.5, -1) @ v
        torch.testing.assert_close(out[s:s + n].transpose(0, 1).float(), ref, atol=2e-2, rtol=2e-2)



@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("seq", [128, 200, 333])
def test_fmha_bwd_fixed_length(cuda_dev, d, causal, seq):
    X = _fmha()
    torch.manual_seed(0)
    b, h = 2, 3
    qkv = torch.randn(b * seq, 3, h, d, device=cuda_dev, dtype=torch.bfloat16, requires_grad=True)
    out = X.FmhaFunc.apply(qkv[:, 0], qkv[:, 1], qkv[:, 2], None, None, None, None, b, causal, None)
    dout = torch.randn_like(out)
    (grad,) = torch.autograd.grad(out, qkv, dout)
    ref_in = qkv.detach().float().requires_grad_()
    q, k, v = (ref_in[:, i].view(b, seq, h, d).transpose(1, 2) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(b * seq, h, d)
    (ref_grad,) = torch.autograd.grad(ref, ref_in, dout.float())
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(grad.float(), ref_grad, atol=5e-2, rtol=5e-2)


def test_fmha_bwd_varlen(cuda_dev):
    X = _fmha()
    torch.manual_seed(0)
    h, d = 2, 64
    lens = [5, 130, 1, 300]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=cuda_dev)
    qkv = torch.randn(sum(lens), 3, h, d, device=cuda_dev, dtype=torch.float16, requires_grad=True)
    out = X.FmhaFunc.apply(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu, cu, max(lens), max(lens), None, False, None)
    dout = torch.randn_like(out)
    (grad,) = torch.autograd.grad(out, qkv, dout)
    ref_in = qkv.detach().float().requires_grad_()
    outs = []
    for i, n in enumerate(lens):
        s = int(cu[i])
        q, k, v = (ref_in[s:s + n, j].transpose(0, 1) for j in range(3))
        outs.append((torch.softmax(q @ k.transpose(1, 2) / d ** 0.5, -1) @ v).transpose(0, 1))
    (ref_grad,) = torch.autograd.grad(torch.cat(outs), ref_in, dout.float())
    torch.testing.assert_close(grad.float(), ref_grad, atol=5e-2, rtol=5e-2)


@pytest.mark.parametrize("world", [2, 8])
def test_nvls_allreduce(cuda_dev, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from apex_b200.parallel import nvls_allreduce as N
    from apex_b200 import _lib
    try:
        _lib.fn("ab_nvls_allreduce")
    except (AttributeError, KeyError, RuntimeError):
        pytest.skip("experimental kernels are not in this build (APEX_B200_EXPERIMENTAL=1 python -m apex_b200._build)")
    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases
    run_distributed(cases.nvls_allreduce_matches_nccl, world, "cuda", backend="nccl")
