"""Conv + bias (+mask) + ReLU and Conv + frozen scale/bias (+residual) + ReLU with hand-written fused pointwise tails and a custom
backward. Reference: apex/contrib/conv_bias_relu/conv_bias_relu.py:9-102 over cuDNN-frontend runtime-fused graphs
(apex/contrib/csrc/conv_bias_relu/conv_bias_relu.cpp, 7 entry points: forward conv+bias+relu graphs and backward drelu+dbias graphs
:1902-1911), and the scale-bias-add-relu / drelu-dscale-dbias graphs of apex/contrib/csrc/bottleneck/bottleneck.cpp:3558-3594.

Here (csrc/conv_epilogue.cu): the convolution is a cuDNN call (a library convolution, as in the reference); everything around it is
ONE kernel per direction —
  forward : out = relu((conv * scale[c]) + bias[c] + residual) * mask, in place over the convolution output (channels-last);
  backward: one pass over (dout, out) produces the ReLU-masked gradient for the residual branch, the scaled gradient that feeds
            cuDNN's dgrad / wgrad, and the per-channel dbias (and dscale) reductions — the eager composition needs four.
Call signatures match the reference: ``ConvBiasReLU(x, weight, bias, padding, stride)`` with bias shaped [1, C, 1, 1], fp16 / bf16
(or fp32) channels-last tensors. CPU tensors take the same math through plain PyTorch ops."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ... import _lib

_lib.declare("ab_conv_epilogue_fwd", "p p p p p p l i i i p")
_lib.declare("ab_conv_epilogue_bwd", "p p p p p p p p p l i i i p")


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _native(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16, torch.float32) and _lib.available()


def _rows_c(t: torch.Tensor):
    """[N, C, H, W] channels-last tensor -> (rows, C) of its [N*H*W, C] memory view."""
    n, c, h, w = t.shape
    return n * h * w, c


def epilogue_fwd(y, scale, bias, z, mask, relu: bool, keep_raw: bool):
    """act(y * scale[c] + bias[c] + z) * mask over a channels-last convolution output ``y`` (in place unless ``keep_raw``).
    scale / bias: fp32 [C] or None; mask: uint8 of y's shape or None. One kernel launch on CUDA; the same math in torch elsewhere."""
    if _native(y):
        rows, C = _rows_c(y)
        out = torch.empty_like(y) if keep_raw else y
        _lib.fn("ab_conv_epilogue_fwd")(y.data_ptr(), _lib.ptr(z), _lib.ptr(mask), _lib.ptr(scale), _lib.ptr(bias), out.data_ptr(), rows, C,
                                        int(relu), _lib.dt(y), _lib.stream_ptr(y.device))
        return out
    o = y.float()
    if scale is not None:
        o = o * scale.view(1, -1, 1, 1)
    if bias is not None:
        o = o + bias.view(1, -1, 1, 1)
    if z is not None:
        o = o + z.float()
    if mask is not None:
        o = o * mask.float()
    return (torch.relu(o) if relu else o).to(y.dtype)


def epilogue_bwd(dout, out, y, mask, scale, relu: bool, want_dz: bool, want_dbias: bool, want_dscale: bool):
    """-> (dy, dz, dbias fp32 [C], dscale fp32 [C]) from ONE pass over (dout, out): g = dout * [out > 0] * mask; dz = g; dy = g * scale;
    dbias = sum g; dscale = sum g * y."""
    C = dout.shape[1]
    if _native(dout):
        rows, _ = _rows_c(dout)
        dy = torch.empty_like(dout)
        dz = (dy if scale is None else torch.empty_like(dout)) if want_dz else None
        dbias = torch.zeros(C, dtype=torch.float32, device=dout.device) if want_dbias else None
        dscale = torch.zeros(C, dtype=torch.float32, device=dout.device) if want_dscale else None
        _lib.fn("ab_conv_epilogue_bwd")(dout.data_ptr(), _lib.ptr(out) if relu else None, _lib.ptr(y) if want_dscale else None, _lib.ptr(mask),
                                        _lib.ptr(scale), dy.data_ptr(), _lib.ptr(dz), _lib.ptr(dbias), _lib.ptr(dscale), rows, C, int(relu),
                                        _lib.dt(dout), _lib.stream_ptr(dout.device))
        return dy, dz, dbias, dscale
    g = dout.float()
    if relu:
        g = g * (out > 0).float()
    if mask is not None:
        g = g * mask.float()
    dbias = g.sum((0, 2, 3)) if want_dbias else None
    dscale = (g * y.float()).sum((0, 2, 3)) if want_dscale else None
    dy = (g * scale.view(1, -1, 1, 1) if scale is not None else g).to(dout.dtype)
    return dy, (g.to(dout.dtype) if want_dz else None), dbias, dscale


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last) if t is not None and t.dim() == 4 else t


class FusedConvEpilogue(torch.autograd.Function):
    """out = act(conv2d(x, weight) * scale + bias + z) * mask  with the fused tail / fused backward described in the module docstring.
    ``scale`` / ``bias``: [C] (any shape with C elements) or None; ``z``: residual of the output's shape or None; ``mask``: 0/1 tensor
    of the output's shape or None; gradients flow to x, weight, bias, scale (when they require grad) and z."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale, z, mask, stride, padding, relu):
        stride, padding = _pair(stride), _pair(padding)
        x, weight = _cl(x), _cl(weight)
        y = _cl(torch.ops.aten.convolution(x, weight.to(x.dtype), None, stride, padding, (1, 1), False, (0, 0), 1))
        need_dscale = scale is not None and scale.requires_grad
        sc32 = scale.detach().reshape(-1).float().contiguous() if scale is not None else None
        b32 = bias.detach().reshape(-1).float().contiguous() if bias is not None else None
        m8 = _cl(mask).to(torch.uint8) if mask is not None else None
        out = epilogue_fwd(y, sc32, b32, _cl(z), m8, bool(relu), need_dscale)   # dscale needs the raw convolution output in the backward
        ctx.save_for_backward(x, weight, out if relu else None, y if need_dscale else None, sc32, m8)
        ctx.cfg = (stride, padding, bool(relu), None if bias is None else (bias.shape, bias.dtype),
                   None if scale is None else (scale.shape, scale.dtype), z is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, out, y, sc32, m8 = ctx.saved_tensors
        stride, padding, relu, bias_meta, scale_meta, has_z = ctx.cfg
        need = ctx.needs_input_grad
        dy, dz, dbias, dscale = epilogue_bwd(_cl(dout), out, y, m8, sc32, relu, has_z and need[4], bias_meta is not None and need[2],
                                             y is not None and need[3])
        dx, dw, _ = torch.ops.aten.convolution_backward(dy, x, weight.to(x.dtype), None, stride, padding, (1, 1), False, (0, 0), 1,
                                                        (bool(need[0]), bool(need[1]), False))
        gb = dbias.reshape(bias_meta[0]).to(bias_meta[1]) if dbias is not None else None
        gs = dscale.reshape(scale_meta[0]).to(scale_meta[1]) if dscale is not None else None
        return dx, (dw.to(weight.dtype) if dw is not None else None), gb, gs, dz, None, None, None, None


def fused_conv_epilogue(x, weight, bias=None, scale=None, z=None, mask=None, stride=1, padding=0, relu=True):
    """Public functional form (used by the four reference entry points below and by contrib.bottleneck)."""
    if x.is_cuda and x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16, torch.float32) and not _lib.available():
        raise _lib.gpu_required_error("conv_bias_relu")
    if x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16, torch.float32):
        return FusedConvEpilogue.apply(x, weight, bias, scale, z, mask, stride, padding, relu)
    y = F.conv2d(x, weight.to(x.dtype), None, stride, padding)   # exotic dtypes (fp64): plain composition
    if scale is not None:
        y = y * scale.reshape(1, -1, 1, 1).to(y.dtype)
    if bias is not None:
        y = y + bias.reshape(1, -1, 1, 1).to(y.dtype)
    if z is not None:
        y = y + z
    if mask is not None:
        y = y * mask.to(y.dtype)
    return F.relu(y) if relu else y


def ConvBiasReLU(x, weight, bias, padding, stride):
    return fused_conv_epilogue(x, weight, bias=bias, stride=stride, padding=padding, relu=True)


def ConvBiasMaskReLU(x, weight, bias, mask, padding, stride):
    return fused_conv_epilogue(x, weight, bias=bias, mask=mask, stride=stride, padding=padding, relu=True)


def ConvBias(x, weight, bias, padding, stride):
    return fused_conv_epilogue(x, weight, bias=bias, stride=stride, padding=padding, relu=False)


def ConvFrozenScaleBiasReLU(x, weight, scale, bias, padding, stride):
    return fused_conv_epilogue(x, weight, bias=bias, scale=scale, stride=stride, padding=padding, relu=True)


class ConvBiasReLU_:
    """``<Name>_.apply`` spelling of the reference (:99-102); each forwards to :class:`FusedConvEpilogue`."""

    apply = staticmethod(ConvBiasReLU)


class ConvBiasMaskReLU_:
    apply = staticmethod(ConvBiasMaskReLU)


class ConvBias_:
    apply = staticmethod(ConvBias)


class ConvFrozenScaleBiasReLU_:
    apply = staticmethod(ConvFrozenScaleBiasReLU)
