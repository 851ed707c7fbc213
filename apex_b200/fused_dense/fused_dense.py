"""FusedDense / FusedDenseGeluDense on the tcgen05 GEMM (csrc/gemm_sm100.cu).

Reference: apex/fused_dense/fused_dense.py:1-114 over csrc/fused_dense_cuda.cu (every FLOP a cuBLASLt call with an epilogue).
Same autograd structure and saved tensors (FusedDenseGeluDense saves ``gelu_in`` and ``output1``); inputs may have any number
of leading dimensions (the reference requires 2-D); the bias gradient is a deterministic column reduction.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..ops import gemm as G


def _cast_if_autocast_enabled(*args):
    if not torch.is_autocast_enabled():
        return args
    dt = torch.get_autocast_dtype("cuda")
    return tuple(a.to(dt) if (torch.is_tensor(a) and a.is_floating_point() and a.is_cuda) else a for a in args)


def _2d(x):
    return x.reshape(-1, x.shape[-1]).contiguous() if (x.dim() != 2 or not x.is_contiguous()) else x


class FusedDenseFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weight, bias):
        x = _2d(input)
        ctx.save_for_backward(x, weight)
        ctx.in_shape = input.shape
        y = G.linear_fwd(x, weight.contiguous(), bias)
        return y.view(*input.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, grad_output):
        x, weight = ctx.saved_tensors
        dy = _2d(grad_output)
        dx = G.linear_dgrad(dy, weight.contiguous()) if ctx.needs_input_grad[0] else None
        dw = G.linear_wgrad(dy, x) if ctx.needs_input_grad[1] else None
        db = G.colsum(dy) if ctx.needs_input_grad[2] else None
        return (dx.view(ctx.in_shape) if dx is not None else None), dw, db


class DenseNoBiasFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weight):
        x = _2d(input)
        ctx.save_for_backward(x, weight)
        ctx.in_shape = input.shape
        return G.linear_fwd(x, weight.contiguous(), None).view(*input.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, grad_output):
        x, weight = ctx.saved_tensors
        dy = _2d(grad_output)
        dx = G.linear_dgrad(dy, weight.contiguous()) if ctx.needs_input_grad[0] else None
        dw = G.linear_wgrad(dy, x) if ctx.needs_input_grad[1] else None
        return (dx.view(ctx.in_shape) if dx is not None else None), dw


class FusedDenseGeluDenseFunc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weight1, bias1, weight2, bias2):
        x = _2d(input)
        gelu_in = torch.empty(x.shape[0], weight1.shape[0], dtype=x.dtype, device=x.device)
        output1 = G.linear_fwd(x, weight1.contiguous(), bias1, epi=G.EPI_BIAS_GELU, aux=gelu_in)   # GEMM1 + bias + GELU (+aux)
        output2 = G.linear_fwd(output1, weight2.contiguous(), bias2)                                 # GEMM2 + bias
        ctx.save_for_backward(x, weight1, weight2, gelu_in, output1)
        ctx.in_shape = input.shape
        return output2.view(*input.shape[:-1], weight2.shape[0])

    @staticmethod
    def backward(ctx, grad_output):
        x, weight1, weight2, gelu_in, output1 = ctx.saved_tensors
        dy = _2d(grad_output)
        dw2 = G.linear_wgrad(dy, output1)
        db2 = G.colsum(dy)
        # dgrad2 fused with gelu' AND with the bias-1 gradient (column sums out of the same epilogue registers)
        d_gelu_in, db1 = G.linear_dgrad(dy, weight2.contiguous(), dgelu_aux=gelu_in, want_colsum=True)
        dw1 = G.linear_wgrad(d_gelu_in, x)
        dx = G.linear_dgrad(d_gelu_in, weight1.contiguous()) if ctx.needs_input_grad[0] else None
        return (dx.view(ctx.in_shape) if dx is not None else None), dw1, db1, dw2, db2


def fused_dense_function(input, weight, bias=None):
    if bias is None:
        args = _cast_if_autocast_enabled(input, weight)
        with torch.amp.autocast("cuda", enabled=False):
            return DenseNoBiasFunc.apply(*args)
    args = _cast_if_autocast_enabled(input, weight, bias)
    with torch.amp.autocast("cuda", enabled=False):
        return FusedDenseFunc.apply(*args)


def fused_dense_gelu_dense_function(input, weight1, bias1, weight2, bias2):
    args = _cast_if_autocast_enabled(input, weight1, bias1, weight2, bias2)
    with torch.amp.autocast("cuda", enabled=False):
        return FusedDenseGeluDenseFunc.apply(*args)


class FusedDenseFP8Func(torch.autograd.Function):
    """Linear with fp8 GEMMs on the tcgen05 kind::f8f6f4 path (per-tensor dynamic scales). Forward: E4M3 x E4M3. ``fp8_backward=True``
    also runs dgrad and wgrad in fp8 (E5M2 gradient x E4M3 weight / activation): the activation is saved as its TRANSPOSED fp8 copy (half the
    bytes of the 16-bit tensor) produced by the same pass that quantises it for the forward. Default: 16-bit dgrad / wgrad.
    (BASELINE.md row 3 'bf16 / fp8 FFN block' -- the reference has no fp8 path at all.)"""

    @staticmethod
    def forward(ctx, input, weight, bias, fp8_backward=False):
        x = _2d(input)
        ctx.in_shape, ctx.has_bias = input.shape, bias is not None
        ctx.fp8_bwd = bool(fp8_backward) and x.is_cuda and x.shape[0] % 16 == 0 and x.shape[1] % 16 == 0 and weight.shape[0] % 16 == 0
        if ctx.fp8_bwd:
            x8, xt8, sx = G.quantize_fp8_dual(x)
            w8, sw = G._quantize_weight_cached(weight.contiguous())
            y = G.gemm_fp8(x8, w8, 1.0, scale_a=sx, scale_b=sw, out_dtype=x.dtype, epi=G.EPI_BIAS if bias is not None else G.EPI_NONE, bias=bias)
            if y is not None:
                ctx.save_for_backward(xt8, sx, weight)
                return y.view(*input.shape[:-1], weight.shape[0])
            ctx.fp8_bwd = False
        ctx.save_for_backward(x, weight)
        y = G.linear_fwd_fp8(x, weight.contiguous(), bias)
        return y.view(*input.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, grad_output):
        dy = _2d(grad_output)
        if ctx.fp8_bwd:
            xt8, sx, weight = ctx.saved_tensors
            r = G.linear_bwd_fp8(dy.contiguous(), weight.contiguous(), xt8, sx, need_dx=ctx.needs_input_grad[0])
            if r is None:
                raise RuntimeError("fp8 backward: shapes must be multiples of 16 (checked in forward)")
            dx, dw = r
            db = G.colsum(dy) if ctx.has_bias else None
            return (dx.view(ctx.in_shape) if dx is not None else None), dw, db, None
        x, weight = ctx.saved_tensors
        dx = G.linear_dgrad(dy, weight.contiguous()) if ctx.needs_input_grad[0] else None
        dw = G.linear_wgrad(dy, x)
        db = G.colsum(dy) if ctx.has_bias else None
        return (dx.view(ctx.in_shape) if dx is not None else None), dw, db, None


class FusedDenseGeluDenseFP8Func(FusedDenseGeluDenseFunc):
    """Dense -> GELU -> Dense with both forward GEMMs on the fp8 path (bias + GELU (+aux) and bias epilogues fused as in the 16-bit
    version). ``fp8_backward=False``: the backward is inherited (16-bit dgrad / wgrad on the saved activations). ``fp8_backward=True``: all
    four backward GEMMs run in fp8 (E5M2 gradients x E4M3 weights / activations; dgrad-2 keeps its fused gelu' epilogue), the block input and
    the hidden activation are saved as transposed fp8 copies made by the forward's own quantisation passes."""

    @staticmethod
    def forward(ctx, input, weight1, bias1, weight2, bias2, fp8_backward=False):
        x = _2d(input)
        gelu_in = torch.empty(x.shape[0], weight1.shape[0], dtype=x.dtype, device=x.device)
        ctx.in_shape = input.shape
        M, K, H, N = x.shape[0], x.shape[1], weight1.shape[0], weight2.shape[0]
        ctx.fp8_bwd = bool(fp8_backward) and x.is_cuda and all(v % 16 == 0 for v in (M, K, H, N))
        if ctx.fp8_bwd:
            x8, xt8, sx = G.quantize_fp8_dual(x)
            w18, sw1 = G._quantize_weight_cached(weight1.contiguous())
            output1 = G.gemm_fp8(x8, w18, 1.0, scale_a=sx, scale_b=sw1, out_dtype=x.dtype, epi=G.EPI_BIAS_GELU, bias=bias1, aux=gelu_in)
            if output1 is not None:
                h8, ht8, sh = G.quantize_fp8_dual(output1)
                w28, sw2 = G._quantize_weight_cached(weight2.contiguous())
                output2 = G.gemm_fp8(h8, w28, 1.0, scale_a=sh, scale_b=sw2, out_dtype=x.dtype, epi=G.EPI_BIAS, bias=bias2)
                if output2 is not None:
                    ctx.save_for_backward(xt8, sx, ht8, sh, gelu_in, weight1, weight2)
                    return output2.view(*input.shape[:-1], N)
            ctx.fp8_bwd = False
        output1 = G.linear_fwd_fp8(x, weight1.contiguous(), bias1, epi=G.EPI_BIAS_GELU, aux=gelu_in)
        output2 = G.linear_fwd_fp8(output1, weight2.contiguous(), bias2)
        ctx.save_for_backward(x, weight1, weight2, gelu_in, output1)
        return output2.view(*input.shape[:-1], weight2.shape[0])

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.fp8_bwd:
            return FusedDenseGeluDenseFunc.backward(ctx, grad_output) + (None,)
        xt8, sx, ht8, sh, gelu_in, weight1, weight2 = ctx.saved_tensors
        dy = _2d(grad_output).contiguous()
        db2 = G.colsum(dy)
        dy8, dyt8, sdy = G.quantize_fp8_dual(dy, torch.float8_e5m2)
        w2t8, sw2 = G._quantize_weight_t_cached(weight2.contiguous())
        # dgrad 2 fused with gelu'(aux); wgrad 2 on the transposed copies
        d_gelu_in = G.gemm_fp8(dy8, w2t8, 1.0, scale_a=sdy, scale_b=sw2, out_dtype=dy.dtype, epi=G.EPI_DGELU, aux=gelu_in)
        dw2 = G.gemm_fp8(dyt8, ht8, 1.0, scale_a=sdy, scale_b=sh, out_dtype=weight2.dtype)
        db1 = G.colsum(d_gelu_in)
        r = G.linear_bwd_fp8(d_gelu_in, weight1.contiguous(), xt8, sx, need_dx=ctx.needs_input_grad[0])
        dx, dw1 = r
        return (dx.view(ctx.in_shape) if dx is not None else None), dw1, db1, dw2, db2, None


def fused_dense_gelu_dense_fp8_function(input, weight1, bias1, weight2, bias2, fp8_backward=False):
    args = _cast_if_autocast_enabled(input, weight1, bias1, weight2, bias2)
    with torch.amp.autocast("cuda", enabled=False):
        return FusedDenseGeluDenseFP8Func.apply(*args, fp8_backward)


def fused_dense_fp8_function(input, weight, bias=None, fp8_backward=False):
    args = _cast_if_autocast_enabled(input, weight, bias)
    with torch.amp.autocast("cuda", enabled=False):
        return FusedDenseFP8Func.apply(*args, fp8_backward)


class FusedDense(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if self.bias is not None:
            bound = 1 / self.in_features ** 0.5
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, input):
        return fused_dense_function(input, self.weight, self.bias)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"


class FusedDenseGeluDense(nn.Module):
    def __init__(self, in_features, intermediate_features, out_features, bias=True):
        super().__init__()
        assert bias, "DenseGeluDense module without bias is currently not supported"
        self.in_features, self.intermediate_features, self.out_features = in_features, intermediate_features, out_features
        self.weight1 = nn.Parameter(torch.empty(intermediate_features, in_features))
        self.bias1 = nn.Parameter(torch.empty(intermediate_features))
        self.weight2 = nn.Parameter(torch.empty(out_features, intermediate_features))
        self.bias2 = nn.Parameter(torch.empty(out_features))
        for w, b, fan in ((self.weight1, self.bias1, in_features), (self.weight2, self.bias2, intermediate_features)):
            nn.init.kaiming_uniform_(w, a=5 ** 0.5)
            nn.init.uniform_(b, -1 / fan ** 0.5, 1 / fan ** 0.5)

    def forward(self, input):
        return fused_dense_gelu_dense_function(input, self.weight1, self.bias1, self.weight2, self.bias2)
