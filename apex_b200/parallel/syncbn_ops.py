"""The ten standalone entry points of the reference ``syncbn`` extension (csrc/syncbn.cpp:71-89, kernels csrc/welford.cu:217-788).

Every op is a phase subset of the one persistent kernel in csrc/syncbn.cu (``phases`` bitmask: 1 = local statistics, 2 = finalize,
4 = elementwise), so the decomposed API and :class:`apex_b200.parallel.SyncBatchNorm` share code. Statistics are fp32; inputs
f32 / f16 / bf16; NCHW (any trailing dims) and channels-last are both native layouts. On CPU tensors a PyTorch oracle runs.
"""
from __future__ import annotations

import torch

from .. import _lib
from .sync_batchnorm import _GroupState, _call, _layout


def _cuda(x):
    return x.is_cuda and _lib.available()


def _bshape(x, nhwc):
    return [1, -1] + [1] * (x.dim() - 2)


def _prep(x, c_last):
    if c_last and x.dim() >= 3 and not x.is_contiguous(memory_format=torch.channels_last if x.dim() == 4 else torch.channels_last_3d):
        # the reference's *_c_last entry points take tensors whose LAST dim is C; present them as logical NCHW views
        raise ValueError("channels-last ops expect a (N, C, ...) tensor in channels_last memory format")
    return _layout(x)


def welford_mean_var(input: torch.Tensor):
    """-> (mean[C], var_biased[C]) fp32 over all dims but 1."""
    x, N, C, HW, nhwc = _layout(input)
    if not _cuda(x):
        xf = x.float().transpose(0, 1).reshape(C, -1)
        return xf.mean(1), xf.var(1, unbiased=False)
    st = _GroupState.get(None, x.device, C)
    mean = torch.empty(C, dtype=torch.float32, device=x.device)
    var = torch.empty(C, dtype=torch.float32, device=x.device)
    _call(st, 0, 3, x, None, None, None, None, N, C, HW, nhwc, None, None, mean, None, var, None, None, 0.0, 0.0, None, None, None, None,
          False, False)
    return mean, var


welford_mean_var_c_last = welford_mean_var


def welford_parallel(mean_all: torch.Tensor, var_all: torch.Tensor, numel: torch.Tensor, eps: float):
    """Chan-merge of per-rank (mean, biased var, count) rows [W, C] -> (mean, var_unbiased, inv_std); counts may differ per rank."""
    n = numel.to(torch.float32).reshape(-1, 1)
    tot = n.sum()
    mean = (mean_all.float() * n).sum(0) / tot
    m2 = (var_all.float() * n).sum(0) + ((mean_all.float() - mean) ** 2 * n).sum(0)
    return mean, m2 / (tot - 1).clamp_min(1.0), torch.rsqrt(m2 / tot + eps)


def batchnorm_forward(input, mean, inv_std, weight=None, shift=None, z=None, fuse_relu=False):
    x, N, C, HW, nhwc = _layout(input)
    if not _cuda(x):
        sh = _bshape(x, nhwc)
        y = (x.float() - mean.view(sh)) * inv_std.view(sh)
        if weight is not None:
            y = y * weight.float().view(sh)
        if shift is not None:
            y = y + shift.float().view(sh)
        if z is not None:
            y = y + z.float()
        return (y.relu() if fuse_relu else y).to(x.dtype)
    st = _GroupState.get(None, x.device, C)
    y = torch.empty_like(x)
    zz = None if z is None else (z.contiguous(memory_format=torch.channels_last) if nhwc and z.dim() == 4 else z.contiguous())
    _call(st, 0, 4, x, None, zz, y, None, N, C, HW, nhwc, None if weight is None else weight.float(), None if shift is None else shift.float(),
          mean, inv_std, None, None, None, 0.0, 0.0, None, None, None, None, fuse_relu, False)
    return y


def batchnorm_forward_c_last(input, z, mean, inv_std, weight, shift, fuse_relu):
    return batchnorm_forward(input, mean, inv_std, weight, shift, z=z, fuse_relu=fuse_relu)


def reduce_bn(grad_output, input, mean, inv_std, weight=None):
    """-> (sum_dy[C], sum_dy_xmu[C], grad_weight[C], grad_bias[C]) (local sums; all-reduce the first two across ranks)."""
    x, N, C, HW, nhwc = _layout(input)
    dy = grad_output.contiguous(memory_format=torch.channels_last) if (nhwc and grad_output.dim() == 4) else grad_output.contiguous()
    if not _cuda(x):
        sh = _bshape(x, nhwc)
        dims = [d for d in range(x.dim()) if d != 1]
        g = dy.float()
        sum_dy = g.sum(dims)
        sum_dy_xmu = (g * (x.float() - mean.view(sh))).sum(dims)
        return sum_dy, sum_dy_xmu, sum_dy_xmu * inv_std, sum_dy
    st = _GroupState.get(None, x.device, C)
    f = lambda: torch.empty(C, dtype=torch.float32, device=x.device)
    sdy, sdx, gw, gb = f(), f(), f(), f()
    _call(st, 1, 3, x, dy, None, None, None, N, C, HW, nhwc, None if weight is None else weight.float(), None, mean, inv_std, None, None, None,
          0.0, 0.0, gw, gb, sdy, sdx, False, False)
    return sdy, sdx, gw, gb


def reduce_bn_c_last(grad_output, input, mean, inv_std, weight=None):
    return reduce_bn(grad_output, input, mean, inv_std, weight)


def batchnorm_backward(grad_output, input, mean, inv_std, weight, sum_dy, sum_dy_xmu, count):
    """dx = (dy - sum_dy/N - (x - mean) * inv_std^2 * sum_dy_xmu/N) * w * inv_std with N = sum(count) over ranks."""
    x, N, C, HW, nhwc = _layout(input)
    dy = grad_output.contiguous(memory_format=torch.channels_last) if (nhwc and grad_output.dim() == 4) else grad_output.contiguous()
    tot = count.to(torch.float32).sum().reshape(1) if torch.is_tensor(count) else torch.tensor([float(count)], device=x.device)
    if not _cuda(x):
        sh = _bshape(x, nhwc)
        w = weight.float().view(sh) if weight is not None else 1.0
        g = dy.float()
        dx = (g - sum_dy.view(sh) / tot - (x.float() - mean.view(sh)) * inv_std.view(sh) ** 2 * sum_dy_xmu.view(sh) / tot) * w * inv_std.view(sh)
        return dx.to(x.dtype)
    st = _GroupState.get(None, x.device, C)
    st.count.copy_(tot)
    dx = torch.empty_like(x)
    _call(st, 1, 4, x, dy, None, dx, None, N, C, HW, nhwc, None if weight is None else weight.float(), None, mean, inv_std, None, None, None,
          0.0, 0.0, None, None, sum_dy.float(), sum_dy_xmu.float(), False, False)
    return dx


def batchnorm_backward_c_last(grad_output, input, mean, inv_std, weight, sum_dy, sum_dy_xmu, count):
    return batchnorm_backward(grad_output, input, mean, inv_std, weight, sum_dy, sum_dy_xmu, count)


def relu_bw_c_last(grad_output, input, z, mean, inv_std, weight, shift):
    """Gradient mask of the fused (bn + add + relu) block: dy where bn(x) + z > 0 else 0 (reference welford.cu:565-604)."""
    y = batchnorm_forward(input, mean, inv_std, weight, shift, z=z, fuse_relu=False)
    return torch.where(y > 0, grad_output, torch.zeros_like(grad_output))
