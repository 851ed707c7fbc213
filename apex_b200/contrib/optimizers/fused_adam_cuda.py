"""The seven entry points of the reference's deprecated ``fused_adam_cuda`` extension (apex/contrib/csrc/optimizers/fused_adam_cuda.cpp:92-104,
kernels fused_adam_cuda_kernel.cu): single-tensor Adam with explicit gradient scale and low-precision copy-out, the *reversible* step
with its conditional undo (a step taken on gradients that turn out to contain inf / nan can be rolled back instead of being skipped
up front), the strided finite check, and the dtype-cast helpers (``maybe_cast_mt`` is what the reference's DistributedFusedAdam uses
to unpack gathered parameters). Tensor programs over PyTorch ops / the multi-tensor engine — same semantics on CPU and CUDA.

mode 0: denominator sqrt(v + eps); mode 1: sqrt(v) + eps (reference kernel ``adamMode_t``)."""
from __future__ import annotations

import math

import torch

from ...ops import amp_C


def _step_size(lr, beta1, beta2, step, bias_correction):
    return lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step) if bias_correction else lr


def _denom(v, eps, mode):
    return (v + eps).sqrt() if mode == 0 else v.sqrt() + eps


@torch.no_grad()
def adam(p, p_copy, m, v, g, lr, beta1, beta2, eps, grad_scale, step, mode, bias_correction, decay):
    sg = g.float() / grad_scale
    m.mul_(beta1).add_(sg, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(sg, sg, value=1 - beta2)
    update = m / _denom(v, eps, mode) + decay * p.float()
    p.add_(update.to(p.dtype), alpha=-_step_size(lr, beta1, beta2, step, bias_correction))
    if p_copy is not None and p_copy.numel():
        p_copy.copy_(p)


reversible_adam = adam  # the forward step is identical; reversibility is a property of maybe_adam_undo


@torch.no_grad()
def maybe_adam_undo(overflow_flag, p, m, v, g, lr, beta1, beta2, eps, grad_scale, step, mode, bias_correction, decay):
    """Roll the step back when ``overflow_flag`` is set: p, m, v return to their values before :func:`reversible_adam`."""
    if not bool(overflow_flag.reshape(-1)[0] != 0):
        return
    ss = _step_size(lr, beta1, beta2, step, bias_correction)
    sg = g.float() / grad_scale
    p_old = (p.float() + ss * (m / _denom(v, eps, mode))) / (1 - ss * decay)
    p.copy_(p_old.to(p.dtype))
    m.sub_(sg, alpha=1 - beta1).div_(beta1)
    v.addcmul_(sg, sg, value=-(1 - beta2)).div_(beta2)


@torch.no_grad()
def strided_check_finite(overflow_flag, p_copy, stride, clear_overflow_first):
    """Sets ``overflow_flag`` if any of every ``stride``-th element of ``p_copy`` is non-finite (a cheap sampled check)."""
    if clear_overflow_first:
        overflow_flag.zero_()
    if not bool(torch.isfinite(p_copy.reshape(-1)[::max(int(stride), 1)].float()).all()):
        overflow_flag.fill_(1)


def adam_mt(chunk_size, overflow_flag, tensor_lists, lr, beta1, beta2, eps, grad_scale, step, mode, bias_correction, decay):
    """tensor_lists = [p, m, v, g] or [p, m, v, g, p_copy] (reference order)."""
    ps, ms, vs, gs = tensor_lists[:4]
    if grad_scale != 1.0:
        gs = [g.float() / grad_scale for g in gs]
    # the multi-tensor kernel applies the bias correction itself; mode there: 0 = L2 (decay added to the gradient), 1 = decoupled
    amp_C.multi_tensor_adam(chunk_size, overflow_flag, [gs, ps, ms, vs], lr, beta1, beta2, eps, step, 1, int(bias_correction), decay)
    if len(tensor_lists) > 4:
        amp_C.multi_tensor_cast(chunk_size, overflow_flag, [ps, tensor_lists[4]])


@torch.no_grad()
def maybe_cast(overflow_flag, p_in, p_out):
    if overflow_flag is None or not bool(overflow_flag.reshape(-1)[0] != 0):
        p_out.copy_(p_in.view(torch.float8_e5m2).float() if p_in.dtype == torch.uint8 else p_in)


def maybe_cast_mt(chunk_size, overflow_flag, tensor_lists):
    if overflow_flag is not None and bool(overflow_flag.reshape(-1)[0] != 0):
        return
    ins = [t.view(torch.float8_e5m2) if t.dtype == torch.uint8 else t for t in tensor_lists[0]]
    amp_C.multi_tensor_cast(chunk_size, None, [ins, tensor_lists[1]])
