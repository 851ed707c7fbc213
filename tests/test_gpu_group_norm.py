"""NHWC GroupNorm(+SiLU) kernel vs torch.nn.functional.group_norm in fp32 (reference test: apex/contrib/test/group_norm)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,C,H,W,G", [(2, 320, 64, 64, 32), (2, 1280, 8, 8, 32), (1, 640, 16, 16, 16), (3, 96, 7, 5, 8), (2, 128, 32, 32, 32),
                                       (2, 2560, 16, 16, 32), (1, 42 * 4, 9, 9, 4)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("act", ["", "silu"])
def test_group_norm(cuda_dev, N, C, H, W, G, dtype, act):
    from apex_b200.contrib.group_norm import GroupNorm
    torch.manual_seed(0)
    gn = GroupNorm(G, C, act=act).to(cuda_dev, dtype)
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5)
        gn.bias.uniform_(-0.5, 0.5)
    x = (torch.randn(N, C, H, W, device=cuda_dev) * 2 + 1).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xr = x.detach().float().requires_grad_(True)
    wr, br = gn.weight.detach().float().requires_grad_(True), gn.bias.detach().float().requires_grad_(True)
    yr = F.group_norm(xr, G, wr, br, 1e-5)
    if act:
        yr = F.silu(yr)
    y = gn(x)
    assert y.is_contiguous(memory_format=torch.channels_last)
    ft, bt = (1e-4, 1e-3) if dtype == torch.float32 else (3e-2, 6e-2)
    torch.testing.assert_close(y.float(), yr, atol=ft, rtol=ft)
    dy = torch.randn_like(y)
    y.backward(dy)
    yr.backward(dy.float())
    torch.testing.assert_close(x.grad.float(), xr.grad, atol=bt, rtol=bt)
    sc = (N * H * W) ** 0.5
    torch.testing.assert_close(gn.weight.grad.float(), wr.grad, atol=bt * sc, rtol=bt)
    torch.testing.assert_close(gn.bias.grad.float(), br.grad, atol=bt * sc, rtol=bt)


@pytest.mark.parametrize("N,C,H,W,G", [(2, 320, 64, 64, 32), (1, 640, 16, 16, 16), (3, 96, 7, 5, 8), (2, 2560, 16, 16, 32), (1, 42 * 4, 9, 9, 4),
                                       (2, 960, 64, 64, 16)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("act", ["", "silu"])
def test_group_norm_streaming_path(cuda_dev, N, C, H, W, G, dtype, act, monkeypatch):
    """csrc/group_norm_stream.cu (two streaming passes per direction; the default for large activations) forced on every shape:
    channel vectors that straddle groups (C / G = 60, 42), odd spatial sizes, the scalar fallback (C % 8 != 0 handled by V = 1)."""
    monkeypatch.setenv("APEX_B200_GN_STREAM_MIN_MB", "0")
    test_group_norm(cuda_dev, N, C, H, W, G, dtype, act)
