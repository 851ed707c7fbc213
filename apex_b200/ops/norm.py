"""Row normalisation ops (LayerNorm / RMSNorm) — python front-end of csrc/layer_norm_{fwd,bwd}.cu.

Reference front-end: csrc/layer_norm_cuda.cpp:8-270 (flatten to [n1, n2], allocate y / mean / invvar, dispatch on dtypes).
"""
from __future__ import annotations

import math

import torch

from .. import _lib

_lib.declare("ab_layer_norm_fwd", "p p p p p p i i f i i i p")
_lib.declare("ab_layer_norm_bwd", "p p p p p p p p p p i i f i i i i p")

_ws: dict = {}


def _flatten(x: torch.Tensor, normalized_shape):
    n2 = math.prod(normalized_shape)
    n1 = x.numel() // n2 if n2 else 0
    return n1, n2


def _workspace(device, n2: int) -> torch.Tensor:
    need = 2 * 148 * 2 * n2
    buf = _ws.get(device)
    if buf is None or buf.numel() < need:
        buf = _ws[device] = torch.empty(need, dtype=torch.float32, device=device)
    return buf


def norm_fwd(x: torch.Tensor, normalized_shape, weight, bias, eps: float, rms: bool, out_dtype=None):
    """-> (y, mean | None, invvar). x is made contiguous; weight/bias (if any) must have the output dtype."""
    if not _lib.available():
        raise _lib.gpu_required_error("fused layer norm")
    x = x.contiguous()
    n1, n2 = _flatten(x, normalized_shape)
    out_dtype = out_dtype or x.dtype
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    invvar = torch.empty(n1, dtype=torch.float32, device=x.device)
    mean = None if rms else torch.empty(n1, dtype=torch.float32, device=x.device)
    if weight is not None and weight.dtype != out_dtype:
        weight = weight.to(out_dtype)
    if bias is not None and bias.dtype != out_dtype:
        bias = bias.to(out_dtype)
    _lib.fn("ab_layer_norm_fwd")(x.data_ptr(), y.data_ptr(), _lib.ptr(mean), invvar.data_ptr(),
                                 _lib.ptr(weight.contiguous() if weight is not None else None),
                                 _lib.ptr(bias.contiguous() if bias is not None else None), n1, n2, float(eps), _lib.dt(x), _lib.dt(y),
                                 int(rms), _lib.stream_ptr(x.device))
    return y, mean, invvar


def norm_bwd(dy: torch.Tensor, saved: torch.Tensor, mean, invvar, normalized_shape, weight, bias, eps: float, rms: bool,
             memory_efficient: bool, in_dtype):
    """-> (dx[in_dtype], dweight | None, dbias | None). ``saved`` is x, or y when memory_efficient."""
    if not _lib.available():
        raise _lib.gpu_required_error("fused layer norm")
    dy = dy.contiguous()
    saved = saved.contiguous()
    n1, n2 = _flatten(dy, normalized_shape)
    out_dtype = dy.dtype
    dx = torch.empty(dy.shape, dtype=in_dtype, device=dy.device)
    dw = db = None
    if weight is not None:
        if weight.dtype != out_dtype:
            weight = weight.to(out_dtype)
        dw = torch.empty(weight.shape, dtype=out_dtype, device=dy.device)
        if bias is not None and not rms:
            if bias.dtype != out_dtype:
                bias = bias.to(out_dtype)
            db = torch.empty(bias.shape, dtype=out_dtype, device=dy.device)
    ws = _workspace(dy.device, n2)
    _lib.fn("ab_layer_norm_bwd")(dy.data_ptr(), saved.data_ptr(), _lib.ptr(mean), invvar.data_ptr(),
                                 _lib.ptr(weight.contiguous() if weight is not None else None),
                                 _lib.ptr(bias.contiguous() if (bias is not None and not rms) else None), dx.data_ptr(), _lib.ptr(dw),
                                 _lib.ptr(db), ws.data_ptr(), n1, n2, float(eps), _lib.DT[in_dtype], _lib.dt(dy), int(rms),
                                 int(memory_efficient), _lib.stream_ptr(dy.device))
    return dx, dw, db
