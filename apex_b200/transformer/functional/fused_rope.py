"""Fused rotary position embedding (csrc/rope.cu): sbhd (+cached cos/sin), packed thd, 2-D. API of the reference extension
``fused_rotary_positional_embedding`` (csrc/megatron/fused_rotary_positional_embedding.cpp:42-193) plus the autograd wrappers
Megatron used to ship (``fused_apply_rotary_pos_emb*``)."""
from __future__ import annotations

import torch

from ... import _lib

_lib.declare("ab_rope", "p p i i i i i i i i i l l l l l l l l l l l l l i i p p p p p p i i i p")


def _native(t):
    return t.is_cuda and _lib.available() and t.dtype in (torch.float16, torch.bfloat16, torch.float32)


def _rotate_half(x):
    x1, x2 = torch.chunk(x, 2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def _ref_apply(t, cos, sin, backward=False):
    r = cos.shape[-1]
    tr, tp = t[..., :r].float(), t[..., r:]
    if not backward:
        o = tr * cos + _rotate_half(tr) * sin
    else:
        o = tr * cos - _rotate_half(tr * sin)  # transpose of the rotation
    return torch.cat((o.to(t.dtype), tp), dim=-1)


def _launch(x, out, mode, is_bwd, cached, n_tokens, s, b, h, d, r, xs, os_, x2d, o2d, ih, iw, freqs, cs, cu, dt_cs):
    _lib.fn("ab_rope")(x.data_ptr(), out.data_ptr(), mode, int(is_bwd), int(cached), n_tokens, s, b, h, d, r, *xs, *os_, *x2d, *o2d, ih, iw,
                       _lib.ptr(freqs), *[_lib.ptr(c) for c in cs], _lib.ptr(cu), (cu.numel() - 1) if cu is not None else 0, dt_cs,
                       _lib.dt(x), _lib.stream_ptr(x.device))


def _sbhd(t, freqs, cos, sin, transpose_output, is_bwd):
    s, b, h, d = t.shape
    cached = cos is not None
    r = (cos if cached else freqs).shape[-1]
    if not _native(t):
        c = cos.float() if cached else torch.cos(freqs.float())
        sn = sin.float() if cached else torch.sin(freqs.float())
        out = _ref_apply(t, c.view(-1, 1, 1, r)[:s], sn.view(-1, 1, 1, r)[:s], is_bwd)
        return out.transpose(0, 1).contiguous().transpose(0, 1) if transpose_output else out   # [s, b, h, d] values in [b, s, h, d] memory
    out = torch.empty((b, s, h, d), dtype=t.dtype, device=t.device).transpose(0, 1) if transpose_output else torch.empty_like(t, memory_format=torch.contiguous_format)
    cs = [cos.contiguous(), sin.contiguous(), None, None] if cached else [None, None, None, None]
    dt_cs = _lib.dt(cos) if cached else 0
    _launch(t, out, 0, is_bwd, cached, s * b, s, b, h, d, r, (t.stride(0), t.stride(1), t.stride(2), t.stride(3)),
            (out.stride(0), out.stride(1), out.stride(2), out.stride(3)), (0, 0, 0), (0, 0), 0, 0,
            None if cached else freqs.float().contiguous(), cs, None, dt_cs)
    return out


class FusedRoPEFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, freqs, transpose_output_memory=False):
        ctx.save_for_backward(freqs)
        ctx.tr = transpose_output_memory
        return _sbhd(t, freqs, None, None, transpose_output_memory, False)

    @staticmethod
    def backward(ctx, grad):
        (freqs,) = ctx.saved_tensors
        return _sbhd(grad, freqs, None, None, ctx.tr, True), None, None


class FusedRoPECachedFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, cos_, sin_, transpose_output_memory=False):
        ctx.save_for_backward(cos_, sin_)
        ctx.tr = transpose_output_memory
        return _sbhd(t, None, cos_, sin_, transpose_output_memory, False)

    @staticmethod
    def backward(ctx, grad):
        cos_, sin_ = ctx.saved_tensors
        return _sbhd(grad, None, cos_, sin_, ctx.tr, True), None, None, None


def _thd(t, cu_seqlens, freqs, is_bwd):
    T, h, d = t.shape
    r = freqs.shape[-1]
    if not _native(t):
        outs = []
        cu = cu_seqlens.tolist()
        for i in range(len(cu) - 1):
            seg = t[cu[i]:cu[i + 1]].unsqueeze(1)
            f = freqs.float().view(-1, 1, 1, r)[:seg.shape[0]]
            outs.append(_ref_apply(seg, torch.cos(f), torch.sin(f), is_bwd).squeeze(1))
        return torch.cat(outs)
    out = torch.empty_like(t, memory_format=torch.contiguous_format)
    cu = cu_seqlens.to(torch.int32).contiguous()
    _launch(t, out, 1, is_bwd, False, T, 0, 1, h, d, r, (t.stride(0), 0, t.stride(1), t.stride(2)),
            (out.stride(0), 0, out.stride(1), out.stride(2)), (0, 0, 0), (0, 0), 0, 0, freqs.float().contiguous(), [None] * 4, cu, 0)
    return out


class FusedRoPETHDFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, cu_seqlens, freqs):
        ctx.save_for_backward(cu_seqlens, freqs)
        return _thd(t, cu_seqlens, freqs, False)

    @staticmethod
    def backward(ctx, grad):
        cu, freqs = ctx.saved_tensors
        return _thd(grad, cu, freqs, True), None, None


def _2d(t, ih, iw, cos_h, sin_h, cos_w, sin_w, is_bwd):
    b, seq, h, d = t.shape
    assert seq == ih * iw
    half = d // 2
    if not _native(t):
        x = t.view(b, ih, iw, h, d)
        ch, sh = cos_h.float().view(1, -1, 1, 1, half)[:, :ih], sin_h.float().view(1, -1, 1, 1, half)[:, :ih]
        cw, sw = cos_w.float().view(1, 1, -1, 1, half)[:, :, :iw], sin_w.float().view(1, 1, -1, 1, half)[:, :, :iw]
        o = torch.cat((_ref_apply(x[..., :half], ch, sh, is_bwd), _ref_apply(x[..., half:], cw, sw, is_bwd)), dim=-1)
        return o.view(b, seq, h, d)
    x5 = t.view(b, ih, iw, h, d)
    out = torch.empty_like(t, memory_format=torch.contiguous_format)
    cs = [c.contiguous() for c in (cos_h, sin_h, cos_w, sin_w)]
    _launch(x5, out, 2, is_bwd, True, b * ih * iw, 0, b, h, d, half, (0, 0, x5.stride(3), x5.stride(4)), (0, 0, out.stride(2), out.stride(3)),
            (x5.stride(0), x5.stride(1), x5.stride(2)), (out.stride(0), out.stride(1)), ih, iw, None, cs, None, _lib.dt(cos_h))
    return out


class FusedRoPE2DFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, img_h, img_w, cos_h, sin_h, cos_w, sin_w):
        ctx.save_for_backward(cos_h, sin_h, cos_w, sin_w)
        ctx.hw = (img_h, img_w)
        return _2d(t, img_h, img_w, cos_h, sin_h, cos_w, sin_w, False)

    @staticmethod
    def backward(ctx, grad):
        cos_h, sin_h, cos_w, sin_w = ctx.saved_tensors
        return _2d(grad.contiguous(), *ctx.hw, cos_h, sin_h, cos_w, sin_w, True), None, None, None, None, None, None


def fused_apply_rotary_pos_emb(t, freqs, transpose_output_memory=False):
    """t [s, b, h, d]; freqs [s, 1, 1, d2] fp32 angles."""
    return FusedRoPEFunc.apply(t, freqs, transpose_output_memory)


def fused_apply_rotary_pos_emb_cached(t, cos_, sin_, transpose_output_memory=False):
    return FusedRoPECachedFunc.apply(t, cos_, sin_, transpose_output_memory)


def fused_apply_rotary_pos_emb_thd(t, cu_seqlens, freqs):
    """t [total_tokens, h, d] packed sequences; cu_seqlens [n_seqs + 1]."""
    return FusedRoPETHDFunc.apply(t, cu_seqlens, freqs)


def fused_apply_rotary_pos_emb_2d(t, img_h, img_w, cos_h, sin_h, cos_w, sin_w):
    """t [b, img_h*img_w, h, d]; first half of d rotates with the row position, second half with the column position."""
    return FusedRoPE2DFunc.apply(t, img_h, img_w, cos_h, sin_h, cos_w, sin_w)
