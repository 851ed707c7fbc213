"""``torch.library`` custom ops for LayerNorm / RMSNorm, so that ``torch.compile`` keeps the fused kernels as opaque graph nodes instead of
tracing into (and replacing) them. Reference: apex/normalization/fused_layer_norm.py:77-330 registers ``apex::fused_layer_norm_affine_fwd / _bwd``
and the RMS twins the same way (``torch.library.custom_op`` + ``register_fake`` + ``register_autograd``).

Two ops serve every flavour (affine / plain, same / mixed dtype, memory-efficient): ``apex_b200::norm_fwd`` and ``apex_b200::norm_bwd``. CUDA
tensors run csrc/layer_norm_{fwd,bwd}.cu; CPU tensors run the explicit formulas below (what the CPU tests exercise). The eager modules keep
their autograd.Function path; these ops are what they switch to while a graph is being compiled."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch

from ..ops import norm as _norm

_HAS = hasattr(torch.library, "custom_op")


def _rows(x: torch.Tensor, normalized_shape) -> int:
    n2 = math.prod(normalized_shape)
    return x.numel() // n2 if n2 else 0


def reference_fwd(x, normalized_shape, weight, bias, eps, rms, out_dtype):
    """(y, mean | None, invvar) with fp32 statistics — the formulas of the kernels, in PyTorch."""
    n1, n2 = _rows(x, normalized_shape), math.prod(normalized_shape)
    xf = x.reshape(n1, n2).float()
    if rms:
        mean = None
        invvar = torch.rsqrt(xf.pow(2).mean(1) + eps)
        xhat = xf * invvar.unsqueeze(1)
    else:
        mean = xf.mean(1)
        invvar = torch.rsqrt(xf.var(1, unbiased=False) + eps)
        xhat = (xf - mean.unsqueeze(1)) * invvar.unsqueeze(1)
    y = xhat
    if weight is not None:
        y = y * weight.reshape(1, n2).float()
    if bias is not None and not rms:
        y = y + bias.reshape(1, n2).float()
    return y.reshape(x.shape).to(out_dtype), mean, invvar


def reference_bwd(dy, saved, mean, invvar, normalized_shape, weight, bias, rms, memory_efficient, in_dtype):
    """(dx, dweight | None, dbias | None); ``saved`` is the input, or the OUTPUT when memory_efficient (x-hat is then recovered from it)."""
    n1, n2 = _rows(dy, normalized_shape), math.prod(normalized_shape)
    g = dy.reshape(n1, n2).float()
    s = saved.reshape(n1, n2).float()
    w = weight.reshape(1, n2).float() if weight is not None else None
    if memory_efficient:
        xhat = s
        if bias is not None and not rms:
            xhat = xhat - bias.reshape(1, n2).float()
        if w is not None:
            xhat = xhat / w
    elif rms:
        xhat = s * invvar.unsqueeze(1)
    else:
        xhat = (s - mean.unsqueeze(1)) * invvar.unsqueeze(1)
    gw = g * w if w is not None else g
    c2 = (gw * xhat).mean(1, keepdim=True)
    dx = gw - xhat * c2
    if not rms:
        dx = dx - gw.mean(1, keepdim=True)
    dx = (dx * invvar.unsqueeze(1)).reshape(dy.shape).to(in_dtype)
    dw = (g * xhat).sum(0).reshape(weight.shape).to(dy.dtype) if weight is not None else None
    db = g.sum(0).reshape(bias.shape).to(dy.dtype) if (bias is not None and not rms and weight is not None) else None
    return dx, dw, db


if _HAS:
    @torch.library.custom_op("apex_b200::norm_fwd", mutates_args=())
    def norm_fwd_op(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], normalized_shape: Sequence[int], eps: float,
                    rms: bool, mixed: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        shape = tuple(normalized_shape)
        out_dtype = weight.dtype if (mixed and weight is not None) else x.dtype
        if x.is_cuda and x.dtype != torch.float64:
            y, mean, invvar = _norm.norm_fwd(x, shape, weight, None if rms else bias, eps, rms, out_dtype)
        else:
            y, mean, invvar = reference_fwd(x, shape, weight, bias, eps, rms, out_dtype)
        return y, (mean if mean is not None else invvar.new_empty(0)), invvar

    @norm_fwd_op.register_fake
    def _(x, weight, bias, normalized_shape, eps, rms, mixed):
        n1 = _rows(x, tuple(normalized_shape))
        out_dtype = weight.dtype if (mixed and weight is not None) else x.dtype
        return (x.new_empty(x.shape, dtype=out_dtype), x.new_empty(0 if rms else n1, dtype=torch.float32), x.new_empty(n1, dtype=torch.float32))

    @torch.library.custom_op("apex_b200::norm_bwd", mutates_args=())
    def norm_bwd_op(dy: torch.Tensor, saved: torch.Tensor, mean: torch.Tensor, invvar: torch.Tensor, weight: Optional[torch.Tensor],
                    bias: Optional[torch.Tensor], normalized_shape: Sequence[int], eps: float, rms: bool, memory_efficient: bool,
                    in_dtype: torch.dtype) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        shape = tuple(normalized_shape)
        m = None if rms else mean
        if dy.is_cuda and dy.dtype != torch.float64:
            dx, dw, db = _norm.norm_bwd(dy, saved, m, invvar, shape, weight, bias, eps, rms, memory_efficient, in_dtype)
        else:
            dx, dw, db = reference_bwd(dy, saved, m, invvar, shape, weight, bias, rms, memory_efficient, in_dtype)
        return dx, (dw if dw is not None else dy.new_empty(0)), (db if db is not None else dy.new_empty(0))   # returns must not alias

    @norm_bwd_op.register_fake
    def _(dy, saved, mean, invvar, weight, bias, normalized_shape, eps, rms, memory_efficient, in_dtype):
        dw = dy.new_empty(weight.shape) if weight is not None else dy.new_empty(0)
        db = dy.new_empty(bias.shape) if (bias is not None and not rms and weight is not None) else dy.new_empty(0)
        return dy.new_empty(dy.shape, dtype=in_dtype), dw, db

    def _setup(ctx, inputs, output):
        x, weight, bias, normalized_shape, eps, rms, mixed = inputs
        y, mean, invvar = output
        ctx.cfg = (tuple(normalized_shape), eps, rms, x.dtype, weight is not None, bias is not None)
        ctx.save_for_backward(x, weight, bias, mean, invvar)

    def _backward(ctx, gy, gmean, ginvvar):
        x, weight, bias, mean, invvar = ctx.saved_tensors
        shape, eps, rms, in_dtype, has_w, has_b = ctx.cfg
        dx, dw, db = norm_bwd_op(gy.contiguous(), x, mean, invvar, weight, bias, shape, eps, rms, False, in_dtype)
        dw = dw.to(weight.dtype) if has_w else None
        db = db.to(bias.dtype) if (has_b and not rms and has_w) else None
        return dx, dw, db, None, None, None, None

    norm_fwd_op.register_autograd(_backward, setup_context=_setup)


def norm(x, weight, bias, normalized_shape, eps, rms=False, mixed=False):
    """Differentiable LayerNorm / RMSNorm through the custom ops (traceable by ``torch.compile`` without a graph break)."""
    if not _HAS:
        raise RuntimeError("torch.library.custom_op is not available in this PyTorch")
    return norm_fwd_op(x.contiguous(), weight, None if rms else bias, list(normalized_shape), float(eps), bool(rms), bool(mixed))[0]
