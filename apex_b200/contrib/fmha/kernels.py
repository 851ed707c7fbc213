"""tcgen05 / TMEM / TMA attention kernels (csrc/fmha_fwd_sm100.cu, csrc/fmha_bwd_sm100.cu) behind a differentiable function.

``FmhaFunc.apply(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, batch, causal, scale, key_bias, dropout_p)``
with q / k / v as [rows, heads, d] views (packed qkv, bshd, sbhd ... any row / head stride that is a multiple of 8 elements),
fp16 / bf16, head dim 64 or 128. Variable-length batches through int32 ``cu_seqlens``; ``key_bias`` [batch, seq_k] fp32 is added to
the scaled scores (key-padding masks); dropout uses a counter-based Philox stream keyed by torch's CUDA generator so the backward
regenerates the forward's mask. Replaces the reference's ``fmhalib`` (apex/contrib/csrc/fmha: sm_80 mma.sync, fp16, d = 64,
seq <= 512) and the cuBLAS-batched-GEMM + softmax/dropout chain of ``fast_multihead_attn``
(apex/contrib/csrc/multihead_attn/multihead_attn_frontend.cpp:573-605)."""
from __future__ import annotations

import torch

from ... import _lib

_lib.declare("ab_fmha_fwd", "p p p p p p p i i i l l i i l l l l l l l l f i p l i f l l i p")
_lib.declare("ab_fmha_bwd", "p p p p p p p p p p p i i i l l i i l l l l l l l l l l l l l l f i p l i f l l i p")


def available() -> bool:
    return _lib.available()


def supported(t: torch.Tensor, head_dim: int) -> bool:
    """Whether the kernels cover this input (otherwise callers compose the generic softmax path)."""
    return t.is_cuda and t.dtype in (torch.float16, torch.bfloat16) and head_dim in (64, 128) and _lib.available()


def next_philox(device, n_calls: int = 1):
    """(seed, offset) for one dropout launch, advancing torch's CUDA generator like a native dropout op would."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + 4 * n_calls)
    return int(seed) & (2**63 - 1), int(offset) & (2**63 - 1)


def _check(t):
    assert t.stride(2) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0, "q/k/v: unit stride along d, row / head strides multiples of 8"


def fmha_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, cu_seqlens_q=None, cu_seqlens_k=None, max_seqlen_q: int | None = None,
             seqlen_k: int | None = None, batch: int | None = None, causal: bool = False, scale: float | None = None, return_lse: bool = False,
             key_bias: torch.Tensor | None = None, dropout_p: float = 0.0, philox=(0, 0), bias_div: int = 0):
    """q [rows_q, heads, d], k / v [rows_k, heads, d]. Fixed-length batches: rows = batch * seqlen; variable length: int32
    ``cu_seqlens`` [batch + 1] on the device. Returns out [rows_q, heads, d] (and the log-sum-exp [rows_q, heads] when requested)."""
    rows_q, heads, d = q.shape
    rows_k = k.shape[0]
    assert q.dtype == k.dtype == v.dtype
    for t in (q, k, v):
        _check(t)
    if cu_seqlens_q is not None:
        batch = cu_seqlens_q.numel() - 1
        assert max_seqlen_q is not None
        cu_seqlens_k = cu_seqlens_q if cu_seqlens_k is None else cu_seqlens_k
        seqlen_k = seqlen_k or max_seqlen_q
    else:
        assert batch is not None and rows_q % batch == 0 and rows_k % batch == 0
        max_seqlen_q, seqlen_k = rows_q // batch, rows_k // batch
    if key_bias is not None:
        key_bias = key_bias.to(torch.float32).contiguous()
        assert key_bias.dim() == 2 and key_bias.shape[1] >= seqlen_k and key_bias.shape[0] == (heads // bias_div if bias_div else batch)
    out = torch.empty(rows_q, heads, d, dtype=q.dtype, device=q.device)
    lse = torch.empty(rows_q, heads, dtype=torch.float32, device=q.device) if return_lse else None
    scale = float(scale if scale is not None else d ** -0.5)
    _lib.fn("ab_fmha_fwd")(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), _lib.ptr(lse), _lib.ptr(cu_seqlens_q), _lib.ptr(cu_seqlens_k),
                           batch, heads, d, rows_q, rows_k, int(max_seqlen_q), int(seqlen_k), q.stride(0), q.stride(1), k.stride(0), k.stride(1),
                           v.stride(0), v.stride(1), out.stride(0), out.stride(1), scale, int(causal), _lib.ptr(key_bias),
                           key_bias.stride(0) if key_bias is not None else 0, int(bias_div), float(dropout_p), int(philox[0]), int(philox[1]), _lib.dt(q),
                           _lib.stream_ptr(q.device))
    return (out, lse) if return_lse else out


def fmha_bwd(dout: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, lse: torch.Tensor, *, cu_seqlens_q=None,
             cu_seqlens_k=None, max_seqlen_q: int | None = None, max_seqlen_k: int | None = None, batch: int | None = None,
             causal: bool = False, scale: float | None = None, key_bias: torch.Tensor | None = None, dropout_p: float = 0.0, philox=(0, 0), bias_div: int = 0):
    """Gradients (dq, dk, dv) of :func:`fmha_fwd` from its output and log-sum-exp (csrc/fmha_bwd_sm100.cu: one kernel instantiation
    for dK / dV, one for dQ, no atomics => deterministic). ``delta = rowsum(dout * out)`` is one small torch reduction."""
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
    if key_bias is not None:
        key_bias = key_bias.to(torch.float32).contiguous()
    delta = (dout.float() * out.float()).sum(-1).contiguous()
    lse = lse.contiguous()
    dq, dk, dv = torch.empty(rows_q, heads, d, dtype=q.dtype, device=q.device), torch.empty_like(k, memory_format=torch.contiguous_format), \
        torch.empty_like(v, memory_format=torch.contiguous_format)
    scale = float(scale if scale is not None else d ** -0.5)
    _lib.fn("ab_fmha_bwd")(q.data_ptr(), k.data_ptr(), v.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                           dv.data_ptr(), _lib.ptr(cu_seqlens_q), _lib.ptr(cu_seqlens_k), batch, heads, d, rows_q, rows_k, int(max_seqlen_q),
                           int(max_seqlen_k), q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1), dout.stride(0),
                           dout.stride(1), dq.stride(0), dq.stride(1), dk.stride(0), dk.stride(1), dv.stride(0), dv.stride(1), scale, int(causal),
                           _lib.ptr(key_bias), key_bias.stride(0) if key_bias is not None else 0, int(bias_div), float(dropout_p), int(philox[0]), int(philox[1]),
                           _lib.dt(q), _lib.stream_ptr(q.device))
    return dq, dk, dv


class FmhaFunc(torch.autograd.Function):
    """Differentiable attention: ``FmhaFunc.apply(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, batch, causal,
    scale, key_bias=None, dropout_p=0.0, bias_div=0)`` with q / k / v as [rows, heads, d] views. ``bias_div`` > 0 selects the bias row by
    ``head // bias_div`` instead of the batch index (used when a [t, b, e] tensor is addressed as one batch of b * heads heads)."""

    @staticmethod
    def forward(ctx, q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, batch, causal, scale, key_bias=None, dropout_p=0.0, bias_div=0):
        philox = next_philox(q.device) if dropout_p > 0.0 else (0, 0)
        out, lse = fmha_fwd(q, k, v, cu_seqlens_q=cu_seqlens_q, cu_seqlens_k=cu_seqlens_k, max_seqlen_q=max_seqlen_q, seqlen_k=max_seqlen_k,
                            batch=batch, causal=causal, scale=scale, return_lse=True, key_bias=key_bias, dropout_p=dropout_p, philox=philox, bias_div=bias_div)
        ctx.save_for_backward(q, k, v, out, lse, cu_seqlens_q, cu_seqlens_k, key_bias)
        ctx.cfg = (max_seqlen_q, max_seqlen_k, batch, causal, scale, dropout_p, philox, bias_div)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse, cu_q, cu_k, key_bias = ctx.saved_tensors
        max_q, max_k, batch, causal, scale, dropout_p, philox, bias_div = ctx.cfg
        dq, dk, dv = fmha_bwd(dout, q, k, v, out, lse, cu_seqlens_q=cu_q, cu_seqlens_k=cu_k, max_seqlen_q=max_q, max_seqlen_k=max_k, batch=batch,
                              causal=causal, scale=scale, key_bias=key_bias, dropout_p=dropout_p, philox=philox, bias_div=bias_div)
        return dq, dk, dv, None, None, None, None, None, None, None, None, None, None


# ---- the kernels' dropout mask in plain torch (tests / debugging): Philox4x32-7 on (q >> 1, k >> 1, b * heads + h, offset) ----
def dropout_keep_mask(batch: int, heads: int, seq_q: int, seq_k: int, p: float, philox, device="cpu") -> torch.Tensor:
    """bool [batch, heads, seq_q, seq_k]: True where the kernels keep the probability (reference implementation of csrc/fmha_common.cuh)."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    seed, offset = philox

    def mul(a: int, x: torch.Tensor):  # 32 x 32 -> (hi, lo) with int64 tensors holding unsigned 32-bit values
        lo16, hi16 = x & 0xFFFF, x >> 16
        a_lo, a_hi = a & 0xFFFF, a >> 16
        ll, lh, hl, hh = a_lo * lo16, a_lo * hi16, a_hi * lo16, a_hi * hi16
        mid = (ll >> 16) + (lh & 0xFFFF) + (hl & 0xFFFF)
        lo = ((mid & 0xFFFF) << 16) | (ll & 0xFFFF)
        hi = hh + (lh >> 16) + (hl >> 16) + (mid >> 16)
        return hi & MASK, lo & MASK

    q2 = torch.arange((seq_q + 1) // 2, dtype=torch.int64, device=device).view(1, -1, 1)
    k2 = torch.arange((seq_k + 1) // 2, dtype=torch.int64, device=device).view(1, 1, -1)
    bh = torch.arange(batch * heads, dtype=torch.int64, device=device).view(-1, 1, 1)
    c0, c1, c2 = (q2 + 0 * k2 + 0 * bh), (k2 + 0 * q2 + 0 * bh), (bh + 0 * q2 + 0 * k2)
    c3 = torch.full_like(c0, offset & MASK)
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    for _ in range(7):
        hi0, lo0 = mul(M0, c0)
        hi1, lo1 = mul(M1, c2)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    thresh = min(int(p * 4294967296.0), 4294967295) if p > 0 else 0
    comps = torch.stack([c0, c1, c2, c3], -1)                      # [bh, q2, k2, 4]: component (q & 1) * 2 + (k & 1)
    full = comps.view(batch * heads, q2.shape[1], k2.shape[2], 2, 2).permute(0, 1, 3, 2, 4).reshape(batch * heads, 2 * q2.shape[1], 2 * k2.shape[2])
    return (full[:, :seq_q, :seq_k] >= thresh).view(batch, heads, seq_q, seq_k)
