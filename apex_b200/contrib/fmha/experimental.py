"""EXPERIMENTAL tcgen05 attention forward (csrc/experimental/fmha_fwd_sm100.cu). Not part of the default build and not yet
validated on hardware: build with ``APEX_B200_EXPERIMENTAL=1 python -m apex_b200._build`` and opt in per call. The supported
:class:`apex_b200.contrib.fmha.FMHA` path is unchanged (varlen packing + SDPA)."""
from __future__ import annotations

import torch

from ... import _lib

_lib.declare("ab_fmha_fwd", "p p p p p p p i i i l l i i l l l l l l l l f i i p")
_lib.declare("ab_fmha_bwd", "p p p p p p p p p p p i i i l l i i l l l l l l l l l l l l l l f i i p")


def available() -> bool:
    if not _lib.available():
        return False
    try:
        _lib.fn("ab_fmha_fwd")
        return True
    except (AttributeError, KeyError):
        return False


def fmha_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, cu_seqlens_q=None, cu_seqlens_k=None, max_seqlen_q: int | None = None,
             seqlen_k: int | None = None, batch: int | None = None, causal: bool = False, scale: float | None = None, return_lse: bool = False):
    """q [rows_q, heads, d], k / v [rows_k, heads, d] (views with arbitrary row / head strides, unit stride along d; fp16 / bf16,
    d in {64, 128}). Fixed-length batches: rows = batch * seqlen; variable length: int32 ``cu_seqlens`` [batch + 1] on the device.
    Returns out [rows_q, heads, d] (and the log-sum-exp [rows_q, heads] when requested)."""
    rows_q, heads, d = q.shape
    rows_k = k.shape[0]
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1 and q.dtype == k.dtype == v.dtype
    if cu_seqlens_q is not None:
        batch = cu_seqlens_q.numel() - 1
        assert max_seqlen_q is not None
        cu_seqlens_k = cu_seqlens_q if cu_seqlens_k is None else cu_seqlens_k
        seqlen_k = seqlen_k or max_seqlen_q
    else:
        assert batch is not None and rows_q % batch == 0 and rows_k % batch == 0
        max_seqlen_q, seqlen_k = rows_q // batch, rows_k // batch
    out = torch.empty(rows_q, heads, d, dtype=q.dtype, device=q.device)
    lse = torch.empty(rows_q, heads, dtype=torch.float32, device=q.device) if return_lse else None
    scale = float(scale if scale is not None else d ** -0.5)
    _lib.fn("ab_fmha_fwd")(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), _lib.ptr(lse), _lib.ptr(cu_seqlens_q), _lib.ptr(cu_seqlens_k),
                           batch, heads, d, rows_q, rows_k, int(max_seqlen_q), int(seqlen_k), q.stride(0), q.stride(1), k.stride(0), k.stride(1),
                           v.stride(0), v.stride(1), out.stride(0), out.stride(1), scale, int(causal), _lib.dt(q), _lib.stream_ptr(q.device))
    return (out, lse) if return_lse else out


def fmha_bwd(dout: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, lse: torch.Tensor, *, cu_seqlens_q=None,
             cu_seqlens_k=None, max_seqlen_q: int | None = None, max_seqlen_k: int | None = None, batch: int | None = None,
             causal: bool = False, scale: float | None = None):
    """Gradients (dq, dk, dv) of :func:`fmha_fwd` from its output and log-sum-exp (csrc/experimental/fmha_bwd_sm100.cu: one kernel
    instantiation for dK / dV, one for dQ, no atomics). ``delta = rowsum(dout * out)`` is one small torch reduction."""
    rows_q, heads, d = q.shape
    rows_k = k.shape[0]
    if cu_seqlens_q is not None:
        batch = cu_seqlens_q.numel() - 1
        assert max_seqlen_q is not None
        cu_seqlens_k = cu_seqlens_q if cu_seqlens_k is None else cu_seqlens_k
        max_seqlen_k = max_seqlen_k or max_seqlen_q
    else:
        assert batch is not None and rows_q % batch == 0 and rows_k % batch == 0
        max_seqlen_q, max_seqlen_k = rows_q // batch, rows_k // batch
    if dout.stride(2) != 1 or dout.stride(0) % 8 or dout.stride(1) % 8:
        dout = dout.contiguous()
    delta = (dout.float() * out.float()).sum(-1).contiguous()
    lse = lse.contiguous()
    dq, dk, dv = torch.empty(rows_q, heads, d, dtype=q.dtype, device=q.device), torch.empty_like(k, memory_format=torch.contiguous_format), \
        torch.empty_like(v, memory_format=torch.contiguous_format)
    scale = float(scale if scale is not None else d ** -0.5)
    _lib.fn("ab_fmha_bwd")(q.data_ptr(), k.data_ptr(), v.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                           dv.data_ptr(), _lib.ptr(cu_seqlens_q), _lib.ptr(cu_seqlens_k), batch, heads, d, rows_q, rows_k, int(max_seqlen_q),
                           int(max_seqlen_k), q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1), dout.stride(0),
                           dout.stride(1), dq.stride(0), dq.stride(1), dk.stride(0), dk.stride(1), dv.stride(0), dv.stride(1), scale, int(causal),
                           _lib.dt(q), _lib.stream_ptr(q.device))
    return dq, dk, dv


class FmhaFunc(torch.autograd.Function):
    """Differentiable attention on the experimental kernels: ``FmhaFunc.apply(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q,
    max_seqlen_k, batch, causal, scale)`` with q / k / v as [rows, heads, d] views."""

    @staticmethod
    def forward(ctx, q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, batch, causal, scale):
        out, lse = fmha_fwd(q, k, v, cu_seqlens_q=cu_seqlens_q, cu_seqlens_k=cu_seqlens_k, max_seqlen_q=max_seqlen_q, seqlen_k=max_seqlen_k,
                            batch=batch, causal=causal, scale=scale, return_lse=True)
        ctx.save_for_backward(q, k, v, out, lse, cu_seqlens_q, cu_seqlens_k)
        ctx.cfg = (max_seqlen_q, max_seqlen_k, batch, causal, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse, cu_q, cu_k = ctx.saved_tensors
        max_q, max_k, batch, causal, scale = ctx.cfg
        dq, dk, dv = fmha_bwd(dout, q, k, v, out, lse, cu_seqlens_q=cu_q, cu_seqlens_k=cu_k, max_seqlen_q=max_q, max_seqlen_k=max_k, batch=batch,
                              causal=causal, scale=scale)
        return dq, dk, dv, None, None, None, None, None, None, None
