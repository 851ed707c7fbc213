"""FusedLayerNorm / FusedRMSNorm (+ "Mixed" variants whose output dtype follows the parameters).

Public surface and behaviour follow the reference apex/normalization/fused_layer_norm.py:670-1037:
  * modules fall back to ``F.layer_norm`` / ``manual_rms_norm`` when tracing, scripting, compiling or on CPU input;
  * functional entry points cast their arguments under autocast and then run with autocast disabled;
  * ``memory_efficient=True`` saves the OUTPUT (+invvar) instead of input+mean and rebuilds x-hat in backward;
  * "mixed dtype": y has the dtype of weight (e.g. bf16 activations with fp32 norm weights -> fp32 output).
Kernels: csrc/layer_norm_fwd.cu, csrc/layer_norm_bwd.cu (one read of x in forward; one read of (dy, x) in backward).
"""
from __future__ import annotations

import numbers

import torch
import torch.nn.functional as F
from torch.nn import init
from torch.nn.parameter import Parameter

from ..ops import norm as _norm


def _cast_if_autocast_enabled(*args):
    if not torch.is_autocast_enabled():
        return args
    dt = torch.get_autocast_dtype("cuda")
    return tuple(a.to(dt) if (torch.is_tensor(a) and a.is_floating_point() and a.is_cuda) else a for a in args)


def manual_rms_norm(input, normalized_shape, weight, eps):
    """Plain PyTorch RMSNorm (CPU / tracing fallback and test oracle), reference :14-35."""
    dims = tuple(i for i in range(-1, -len(normalized_shape) - 1, -1))
    variance = input.float().pow(2).mean(dims, keepdim=True)
    out = input.float() * torch.rsqrt(variance + eps)
    if weight is None:
        return out.to(input.dtype)
    if weight.dtype in (torch.float16, torch.bfloat16):
        out = out.to(weight.dtype)
    return (weight * out).to(weight.dtype if weight.dtype != input.dtype else input.dtype)


class _NormFunction(torch.autograd.Function):
    """One autograd node for all eight (LN|RMS) x (affine|plain) x (same|mixed dtype) flavours."""

    @staticmethod
    def forward(ctx, input, weight, bias, normalized_shape, eps, memory_efficient, rms, mixed):
        ctx.normalized_shape = tuple(normalized_shape)
        ctx.eps, ctx.rms, ctx.memory_efficient = eps, rms, memory_efficient
        ctx.in_dtype = input.dtype
        ctx.has_bias = bias is not None
        out_dtype = weight.dtype if (mixed and weight is not None) else input.dtype
        x = input.contiguous()
        y, mean, invvar = _norm.norm_fwd(x, ctx.normalized_shape, weight, None if rms else bias, eps, rms, out_dtype)
        if memory_efficient:
            ctx.save_for_backward(y, weight, bias, None, invvar)
        else:
            ctx.save_for_backward(x, weight, bias, mean, invvar)
        return y

    @staticmethod
    def backward(ctx, grad_output):
        saved, weight, bias, mean, invvar = ctx.saved_tensors
        dx, dw, db = _norm.norm_bwd(grad_output.contiguous(), saved, mean, invvar, ctx.normalized_shape, weight, bias, ctx.eps,
                                    ctx.rms, ctx.memory_efficient, ctx.in_dtype)
        if dw is not None and dw.dtype != weight.dtype:
            dw = dw.to(weight.dtype)
        if db is not None and bias is not None and db.dtype != bias.dtype:
            db = db.to(bias.dtype)
        return dx, dw, (db if ctx.has_bias else None), None, None, None, None, None


def _shape(normalized_shape):
    if isinstance(normalized_shape, numbers.Integral):
        normalized_shape = (normalized_shape,)
    return tuple(normalized_shape)


def _run(input, weight, bias, normalized_shape, eps, memory_efficient, rms, mixed):
    if not input.is_cuda or input.dtype == torch.float64:  # CPU tensors (reference :815-822) and fp64 (no fp64 kernel): plain PyTorch
        shape = _shape(normalized_shape)
        if rms:
            return manual_rms_norm(input, shape, weight, eps)
        return torch.nn.functional.layer_norm(input, shape, weight, bias, eps)
    if mixed and weight is not None:
        args = _cast_if_autocast_enabled(input) + (weight, bias)
    else:
        args = _cast_if_autocast_enabled(input, weight, bias)
    with torch.amp.autocast("cuda", enabled=False):
        return _NormFunction.apply(args[0], args[1], args[2], _shape(normalized_shape), eps, memory_efficient, rms, mixed)


def fused_layer_norm_affine(input, weight, bias, normalized_shape, eps=1e-6, memory_efficient=False):
    return _run(input, weight, bias, normalized_shape, eps, memory_efficient, False, False)


def fused_layer_norm(input, normalized_shape, eps=1e-6, memory_efficient=False):
    return _run(input, None, None, normalized_shape, eps, memory_efficient, False, False)


def mixed_dtype_fused_layer_norm_affine(input, weight, bias, normalized_shape, eps=1e-6, memory_efficient=False):
    return _run(input, weight, bias, normalized_shape, eps, memory_efficient, False, True)


def fused_rms_norm_affine(input, weight, normalized_shape, eps=1e-6, memory_efficient=False):
    return _run(input, weight, None, normalized_shape, eps, memory_efficient, True, False)


def fused_rms_norm(input, normalized_shape, eps=1e-6, memory_efficient=False):
    return _run(input, None, None, normalized_shape, eps, memory_efficient, True, False)


def mixed_dtype_fused_rms_norm_affine(input, weight, normalized_shape, eps=1e-6, memory_efficient=False):
    return _run(input, weight, None, normalized_shape, eps, memory_efficient, True, True)


class _ApexNamedFunction:
    """The reference exposes one autograd.Function per flavour (fused_layer_norm.py:38-420: FusedLayerNormAffineFunction, ...) and
    downstream code (Megatron-LM) calls ``XFunction.apply(...)`` directly. They all map onto :class:`_NormFunction`."""

    _rms = False
    _mixed = False
    _affine = True
    _has_bias = True

    @classmethod
    def apply(cls, input, *args):
        args = list(args)
        weight = args.pop(0) if cls._affine else None
        bias = args.pop(0) if (cls._affine and cls._has_bias) else None
        normalized_shape, eps = args[0], args[1]
        memory_efficient = args[2] if len(args) > 2 else False
        return _run(input, weight, bias, normalized_shape, eps, memory_efficient, cls._rms, cls._mixed)


class FusedLayerNormAffineFunction(_ApexNamedFunction):
    pass


class FusedLayerNormAffineMixedDtypesFunction(_ApexNamedFunction):
    _mixed = True


class FusedLayerNormFunction(_ApexNamedFunction):
    _affine = False


class FusedRMSNormAffineFunction(_ApexNamedFunction):
    _rms, _has_bias = True, False


class FusedRMSNormAffineMixedDtypesFunction(_ApexNamedFunction):
    _rms, _has_bias, _mixed = True, False, True


class FusedRMSNormFunction(_ApexNamedFunction):
    _rms, _affine = True, False


def supports_custom_op() -> bool:
    return hasattr(torch.library, "custom_op")


def _compiled_cuda(input) -> bool:
    """While ``torch.compile`` traces a CUDA graph the modules call the ``apex_b200::norm_fwd / norm_bwd`` custom ops (custom_ops.py): the fused
    kernels stay in the compiled graph as opaque nodes (what the reference's ``apex::fused_layer_norm_affine_fwd`` ops are for)."""
    return torch.compiler.is_compiling() and input.is_cuda and input.dtype != torch.float64 and supports_custom_op()


def _custom_op_norm(input, weight, bias, normalized_shape, eps, rms, mixed):
    from . import custom_ops

    if mixed or weight is None:
        x = _cast_if_autocast_enabled(input)[0]
    else:
        x, weight, bias = _cast_if_autocast_enabled(input, weight, bias)
    return custom_ops.norm(x, weight, bias, _shape(normalized_shape), eps, rms=rms, mixed=mixed)


def _use_fallback(input) -> bool:
    return (torch.jit.is_tracing() or torch.jit.is_scripting() or torch.compiler.is_compiling() or not input.is_cuda
            or input.dtype == torch.float64)


class FusedLayerNorm(torch.nn.Module):
    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True, memory_efficient=False):
        super().__init__()
        self.normalized_shape = torch.Size(_shape(normalized_shape))
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        self.memory_efficient = memory_efficient
        if elementwise_affine:
            self.weight = Parameter(torch.empty(*self.normalized_shape))
            self.bias = Parameter(torch.empty(*self.normalized_shape))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.elementwise_affine:
            init.ones_(self.weight)
            init.zeros_(self.bias)

    def forward(self, input):
        if _compiled_cuda(input):
            return _custom_op_norm(input, self.weight, self.bias, self.normalized_shape, self.eps, False, False)
        if _use_fallback(input):
            return F.layer_norm(input, self.normalized_shape, self.weight, self.bias, self.eps)
        if self.elementwise_affine:
            return fused_layer_norm_affine(input, self.weight, self.bias, self.normalized_shape, self.eps, self.memory_efficient)
        return fused_layer_norm(input, self.normalized_shape, self.eps, self.memory_efficient)

    def extra_repr(self):
        return "{normalized_shape}, eps={eps}, elementwise_affine={elementwise_affine}".format(**self.__dict__)


class FusedRMSNorm(torch.nn.Module):
    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True, memory_efficient=False):
        super().__init__()
        self.normalized_shape = torch.Size(_shape(normalized_shape))
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        self.memory_efficient = memory_efficient
        if elementwise_affine:
            self.weight = Parameter(torch.empty(*self.normalized_shape))
        else:
            self.register_parameter("weight", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.elementwise_affine:
            init.ones_(self.weight)

    def forward(self, input):
        if _compiled_cuda(input):
            return _custom_op_norm(input, self.weight, None, self.normalized_shape, self.eps, True, False)
        if _use_fallback(input):
            return manual_rms_norm(input, self.normalized_shape, self.weight, self.eps)
        if self.elementwise_affine:
            return fused_rms_norm_affine(input, self.weight, self.normalized_shape, self.eps, self.memory_efficient)
        return fused_rms_norm(input, self.normalized_shape, self.eps, self.memory_efficient)

    def extra_repr(self):
        return "{normalized_shape}, eps={eps}, elementwise_affine={elementwise_affine}".format(**self.__dict__)


class MixedFusedLayerNorm(FusedLayerNorm):
    """Output dtype follows the PARAMETER dtype (Megatron-style fp32 norm params with 16-bit activations)."""

    def __init__(self, normalized_shape, eps=1e-5, *, memory_efficient=False, **kwargs):
        if "elementwise_affine" in kwargs:
            import warnings

            warnings.warn("MixedFusedLayerNorm does not support `elementwise_affine` argument")
            if not kwargs.pop("elementwise_affine"):
                raise RuntimeError("MixedFusedLayerNorm does not support `elementwise_affine = False`")
        super().__init__(normalized_shape=normalized_shape, eps=eps, elementwise_affine=True, memory_efficient=memory_efficient)

    def forward(self, input):
        if _compiled_cuda(input):
            return _custom_op_norm(input, self.weight, self.bias, self.normalized_shape, self.eps, False, True)
        if _use_fallback(input):
            return F.layer_norm(input, self.normalized_shape, self.weight, self.bias, self.eps)
        return mixed_dtype_fused_layer_norm_affine(input, self.weight, self.bias, self.normalized_shape, self.eps, self.memory_efficient)


class MixedFusedRMSNorm(FusedRMSNorm):
    def __init__(self, normalized_shape, eps=1e-5, *, memory_efficient=False, **kwargs):
        if "elementwise_affine" in kwargs:
            import warnings

            warnings.warn("MixedFusedRMSNorm does not support `elementwise_affine` argument")
            if not kwargs.pop("elementwise_affine"):
                raise RuntimeError("MixedFusedRMSNorm does not support `elementwise_affine = False`")
        super().__init__(normalized_shape=normalized_shape, eps=eps, elementwise_affine=True, memory_efficient=memory_efficient)

    def forward(self, input):
        if _compiled_cuda(input):
            return _custom_op_norm(input, self.weight, None, self.normalized_shape, self.eps, True, True)
        if _use_fallback(input):
            return manual_rms_norm(input, self.normalized_shape, self.weight, self.eps)
        return mixed_dtype_fused_rms_norm_affine(input, self.weight, self.normalized_shape, self.eps, self.memory_efficient)
