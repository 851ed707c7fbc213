"""``index_mul_2d(in1, in2, idx1) = in1[idx1] * in2`` fused (no gathered temporary), with first- and second-order gradients.
Reference: apex/contrib/index_mul_2d/index_mul_2d.py:6-135 (fp32/fp16, 2-D, index on dim 0). The double-backward is expressed with
differentiable torch ops (it is off the hot path)."""
from __future__ import annotations

import torch

from ... import _lib

_lib.declare("ab_index_mul_2d_fwd", "p p p p l i i p")
_lib.declare("ab_index_mul_2d_bwd", "p p p p p p l i i p")


def _native(t):
    return t.is_cuda and _lib.available() and t.dtype in (torch.float32, torch.float16, torch.bfloat16)


class _IndexMul2dBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in1, in2, idx1, grad_out):
        ctx.save_for_backward(in1, in2, idx1, grad_out)
        gin1 = torch.zeros(in1.shape, dtype=torch.float32, device=in1.device)
        gin2 = torch.empty_like(in2)
        _lib.fn("ab_index_mul_2d_bwd")(in1.data_ptr(), in2.data_ptr(), idx1.data_ptr(), grad_out.data_ptr(), gin1.data_ptr(), gin2.data_ptr(),
                                       in2.shape[0], in2.shape[1], _lib.dt(in1), _lib.stream_ptr(in1.device))
        return gin1.to(in1.dtype), gin2

    @staticmethod
    def backward(ctx, ggin1, ggin2):
        in1, in2, idx1, grad_out = ctx.saved_tensors
        # d(gin1)/d(in2) = scatter(grad_out), d(gin1)/d(grad_out) = scatter(in2); d(gin2)/d(in1) = grad_out (gathered), d(gin2)/d(grad_out) = in1[idx]
        gg1 = ggin1.index_select(0, idx1)
        g_in1 = torch.zeros_like(in1).index_add_(0, idx1, ggin2 * grad_out)
        g_in2 = gg1 * grad_out
        g_gout = gg1 * in2 + ggin2 * in1.index_select(0, idx1)
        return g_in1, g_in2, None, g_gout


class _IndexMul2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in1, in2, idx1):
        out = torch.empty_like(in2)
        _lib.fn("ab_index_mul_2d_fwd")(in1.data_ptr(), in2.data_ptr(), idx1.data_ptr(), out.data_ptr(), in2.shape[0], in2.shape[1],
                                       _lib.dt(in1), _lib.stream_ptr(in1.device))
        ctx.save_for_backward(in1, in2, idx1)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        in1, in2, idx1 = ctx.saved_tensors
        gin1, gin2 = _IndexMul2dBackward.apply(in1, in2, idx1, grad_out.contiguous())
        return gin1, gin2, None


def index_mul_2d(in1: torch.Tensor, in2: torch.Tensor, idx1: torch.Tensor) -> torch.Tensor:
    assert in2.size(0) == idx1.size(0)
    if in1.dim() != 2 or in2.dim() != 2:
        raise RuntimeError("in1 and in2 must be 2-dimension tensor.")
    if idx1.dim() != 1:
        raise RuntimeError("idx1 must be 1-dimension tensor.")
    if in1.dtype != in2.dtype:
        raise RuntimeError("input1'dtype and input2's dtype must be the same")
    if not _native(in1):
        return in1.index_select(0, idx1) * in2
    return _IndexMul2d.apply(in1.contiguous(), in2.contiguous(), idx1.contiguous().to(torch.int64))


# reference class names (index_mul_2d.py:6-133)
IndexMul2d_ = _IndexMul2d
IndexMul2dBackward_ = _IndexMul2dBackward
