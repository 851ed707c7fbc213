"""Scaled (masked / causal) softmax — python layer over csrc/softmax.cu.

The reference ships four orphaned extensions (scaled_softmax_cuda, scaled_masked_softmax_cuda,
scaled_upper_triang_masked_softmax_cuda, generic_scaled_masked_softmax_cuda; csrc/megatron/*.cpp) whose python wrappers were
removed; the functions and the Megatron ``FusedScaleMaskSoftmax`` module are provided here. No key-length caps (the reference
limits sk to 16384 / 4096) and no ``get_batch_per_block`` divisibility rules."""
from __future__ import annotations

import enum

import torch

from ... import _lib

_lib.declare("ab_softmax_fwd", "p p p f l i i i i i i p")
_lib.declare("ab_softmax_bwd", "p p p f l i i p")


class AttnMaskType(enum.Enum):
    padding = 1
    causal = 2


def _native(x):
    return x.is_cuda and _lib.available() and x.dtype in (torch.float16, torch.bfloat16, torch.float32)


def _ref_forward(x, mask, scale, causal):
    t = x.float() * scale
    if causal:
        sq, sk = x.shape[-2], x.shape[-1]
        cm = torch.triu(torch.ones(sq, sk, dtype=torch.bool, device=x.device), diagonal=1)
        t = t.masked_fill(cm, float("-inf"))
    if mask is not None:
        t = t.masked_fill(mask.bool(), -10000.0)
    y = torch.softmax(t, dim=-1)
    if mask is not None:
        y = y * (~mask.bool().all(dim=-1, keepdim=True))
    return y.to(x.dtype)


def _fwd(x, mask, scale, mode):
    x = x.contiguous()
    sk, sq = x.shape[-1], x.shape[-2]
    rows = x.numel() // sk
    if not _native(x):
        return _ref_forward(x, mask, scale, mode == 2)
    y = torch.empty_like(x)
    heads, per_batch = 1, 0
    m = None
    if mode == 1:
        m = mask.contiguous()
        if m.dtype != torch.uint8:
            m = m.to(torch.uint8)
        heads = x.shape[1] if x.dim() == 4 else 1
        per_batch = 1 if (m.shape[0] != 1) else 0
    _lib.fn("ab_softmax_fwd")(x.data_ptr(), y.data_ptr(), _lib.ptr(m), float(scale), rows, sk, sq, heads, per_batch, mode, _lib.dt(x),
                              _lib.stream_ptr(x.device))
    return y


def _bwd(dy, y, scale):
    dy = dy.contiguous()
    if not _native(y):
        yf, gf = y.float(), dy.float()
        return (scale * yf * (gf - (gf * yf).sum(-1, keepdim=True))).to(y.dtype)
    sk = y.shape[-1]
    dx = torch.empty_like(y)
    _lib.fn("ab_softmax_bwd")(dy.data_ptr(), y.data_ptr(), dx.data_ptr(), float(scale), y.numel() // sk, sk, _lib.dt(y),
                              _lib.stream_ptr(y.device))
    return dx


class ScaledSoftmax(torch.autograd.Function):
    """softmax(scale * x) over the last dim of x[b, h, sq, sk]."""

    @staticmethod
    def forward(ctx, inputs, scale):
        y = _fwd(inputs, None, scale, 0)
        ctx.save_for_backward(y)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, grad):
        (y,) = ctx.saved_tensors
        return _bwd(grad, y, ctx.scale), None


class ScaledMaskedSoftmax(torch.autograd.Function):
    """mask: uint8/bool [b or 1, 1, sq, sk], 1 = masked out (filled with -10000); fully masked rows give zeros."""

    @staticmethod
    def forward(ctx, inputs, mask, scale):
        y = _fwd(inputs, mask, scale, 1)
        ctx.save_for_backward(y)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, grad):
        (y,) = ctx.saved_tensors
        return _bwd(grad, y, ctx.scale), None, None


GenericScaledMaskedSoftmax = ScaledMaskedSoftmax  # one kernel family handles every key length


class ScaledUpperTriangMaskedSoftmax(torch.autograd.Function):
    """Causal softmax on x[attn_batches, sq, sk] (sq == sk): row q attends to keys 0..q, the rest is zero."""

    @staticmethod
    def forward(ctx, inputs, scale):
        y = _fwd(inputs, None, scale, 2)
        ctx.save_for_backward(y)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, grad):
        (y,) = ctx.saved_tensors
        return _bwd(grad, y, ctx.scale), None


def scaled_softmax(x, scale=1.0):
    return ScaledSoftmax.apply(x, scale)


def scaled_masked_softmax(x, mask, scale=1.0):
    return ScaledMaskedSoftmax.apply(x, mask, scale) if mask is not None else ScaledSoftmax.apply(x, scale)


def scaled_upper_triang_masked_softmax(x, scale=1.0):
    return ScaledUpperTriangMaskedSoftmax.apply(x, scale)


class FusedScaleMaskSoftmax(torch.nn.Module):
    """Megatron's fused scale + mask + softmax module (apex.transformer.functional.FusedScaleMaskSoftmax contract)."""

    def __init__(self, input_in_fp16, input_in_bf16, attn_mask_type, scaled_masked_softmax_fusion, mask_func, softmax_in_fp32, scale):
        super().__init__()
        self.input_in_fp16, self.input_in_bf16 = input_in_fp16, input_in_bf16
        if input_in_fp16 and input_in_bf16:
            raise RuntimeError("both fp16 and bf16 flags cannot be active at the same time.")
        self.input_in_float16 = input_in_fp16 or input_in_bf16
        self.attn_mask_type = attn_mask_type
        self.scaled_masked_softmax_fusion = scaled_masked_softmax_fusion
        self.mask_func = mask_func
        self.softmax_in_fp32 = softmax_in_fp32
        self.scale = scale
        if not (self.scale is None or softmax_in_fp32):
            raise RuntimeError("softmax should be in fp32 when scaled")

    def is_kernel_available(self, mask, b, np, sq, sk):
        return bool(self.scaled_masked_softmax_fusion and self.input_in_float16)

    def forward(self, input, mask):
        assert input.dim() == 4
        if self.is_kernel_available(mask, *input.size()):
            return self.forward_fused_softmax(input, mask)
        return self.forward_torch_softmax(input, mask)

    def forward_fused_softmax(self, input, mask):
        scale = self.scale if self.scale is not None else 1.0
        if self.attn_mask_type == AttnMaskType.causal:
            b, np_, sq, sk = input.size()
            assert sq == sk, "causal mask is only for self attention"
            return scaled_upper_triang_masked_softmax(input.view(-1, sq, sk), scale).view(b, np_, sq, sk)
        return scaled_masked_softmax(input, mask, scale)

    def forward_torch_softmax(self, input, mask):
        if self.input_in_float16 and self.softmax_in_fp32:
            input = input.float()
        if self.scale is not None:
            input = input * self.scale
        mask_output = self.mask_func(input, mask) if mask is not None else input
        probs = torch.nn.Softmax(dim=-1)(mask_output)
        if self.input_in_float16 and self.softmax_in_fp32:
            probs = probs.half() if self.input_in_fp16 else probs.bfloat16()
        return probs
