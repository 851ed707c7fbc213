"""Extension-name compatibility: downstream code (Megatron-LM and friends) imports the reference's compiled extensions by name and
calls their raw entry points (``scaled_upper_triang_masked_softmax_cuda.forward``, ``fused_layer_norm_cuda.forward_affine``,
``fused_weight_gradient_mlp_cuda.wgrad_gemm_accum_fp32`` ...). ``install()`` registers modules with those names in ``sys.modules``
whose functions run this library's kernels with the reference's argument order and return conventions
(SURVEY.md §2.12 lists the extensions; signatures from csrc/*.cpp of the reference).
"""
from __future__ import annotations

import sys
import types

import torch


def _mod(name: str, **fns) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__doc__ = f"apex_b200 implementation of the reference extension `{name}`"
    for k, v in fns.items():
        setattr(m, k, v)
    return m


def _softmax_mods():
    from .transformer.functional import fused_softmax as S

    def get_batch_per_block(sq, sk, b, np):  # launch-shape helper of the reference kernels; any value is legal here
        return 1

    return {
        "scaled_softmax_cuda": _mod("scaled_softmax_cuda", forward=lambda x, scale: S._fwd(x, None, scale, 0),
                                    backward=lambda dy, y, scale: S._bwd(dy, y, scale)),
        "scaled_masked_softmax_cuda": _mod("scaled_masked_softmax_cuda", forward=lambda x, mask, scale: S._fwd(x, mask, scale, 1),
                                           backward=lambda dy, y, scale: S._bwd(dy, y, scale), get_batch_per_block=get_batch_per_block),
        "generic_scaled_masked_softmax_cuda": _mod("generic_scaled_masked_softmax_cuda", forward=lambda x, mask, scale: S._fwd(x, mask, scale, 1),
                                                   backward=lambda dy, y, scale: S._bwd(dy, y, scale)),
        "scaled_upper_triang_masked_softmax_cuda": _mod("scaled_upper_triang_masked_softmax_cuda", forward=lambda x, scale: S._fwd(x, None, scale, 2),
                                                        backward=lambda dy, y, scale: S._bwd(dy, y, scale)),
    }


def _layer_norm_mod():
    from .ops import norm as N

    def _cpu(x, shape, w, b, eps, rms):
        xf = x.float()
        dims = tuple(range(x.dim() - len(shape), x.dim()))
        if rms:
            invvar = torch.rsqrt(xf.pow(2).mean(dims, keepdim=True) + eps)
            y, mean = xf * invvar, None
        else:
            mean = xf.mean(dims, keepdim=True)
            invvar = torch.rsqrt(xf.var(dims, unbiased=False, keepdim=True) + eps)
            y = (xf - mean) * invvar
        if w is not None:
            y = y * w.float()
        if b is not None:
            y = y + b.float()
        return y.to(w.dtype if w is not None else x.dtype), (None if mean is None else mean.reshape(-1)), invvar.reshape(-1)

    def _fwd(x, shape, w, b, eps, rms, mixed=False):
        shape = tuple(shape)
        if not x.is_cuda:
            return _cpu(x, shape, w, b, eps, rms)
        return N.norm_fwd(x, shape, w, b, eps, rms, w.dtype if (mixed and w is not None) else None)

    def _bwd(dy, mean, invvar, saved, shape, w, b, eps, rms, memory_efficient):
        return N.norm_bwd(dy, saved, mean, invvar, tuple(shape), w, b, eps, rms, memory_efficient, saved.dtype if not memory_efficient else dy.dtype)

    return _mod(
        "fused_layer_norm_cuda",
        forward=lambda x, shape, eps: _fwd(x, shape, None, None, eps, False),
        forward_affine=lambda x, shape, w, b, eps: _fwd(x, shape, w, b, eps, False),
        forward_affine_mixed_dtypes=lambda x, shape, w, b, eps: _fwd(x, shape, w, b, eps, False, True),
        backward=lambda dy, mean, invvar, saved, shape, eps, memory_efficient=False: _bwd(dy, mean, invvar, saved, shape, None, None, eps, False, memory_efficient)[0],
        backward_affine=lambda dy, mean, invvar, saved, shape, w, b, eps, memory_efficient=False: _bwd(dy, mean, invvar, saved, shape, w, b, eps, False, memory_efficient),
        rms_forward=lambda x, shape, eps: _fwd(x, shape, None, None, eps, True)[::2],
        rms_forward_affine=lambda x, shape, w, eps: _fwd(x, shape, w, None, eps, True)[::2],
        rms_forward_affine_mixed_dtypes=lambda x, shape, w, eps: _fwd(x, shape, w, None, eps, True, True)[::2],
        rms_backward=lambda dy, invvar, saved, shape, eps, memory_efficient=False: _bwd(dy, None, invvar, saved, shape, None, None, eps, True, memory_efficient)[0],
        rms_backward_affine=lambda dy, invvar, saved, shape, w, eps, memory_efficient=False: _bwd(dy, None, invvar, saved, shape, w, None, eps, True, memory_efficient)[:2],
    )


def _rope_mod():
    from .transformer.functional import fused_rope as R

    return _mod(
        "fused_rotary_positional_embedding",
        forward=lambda t, freqs, transpose_output=False: R._sbhd(t, freqs, None, None, transpose_output, False),
        backward=lambda g, freqs, transpose_output=False: R._sbhd(g, freqs, None, None, transpose_output, True),
        forward_cached=lambda t, cos, sin, transpose_output=False: R._sbhd(t, None, cos, sin, transpose_output, False),
        backward_cached=lambda g, cos, sin, transpose_output=False: R._sbhd(g, None, cos, sin, transpose_output, True),
        forward_thd=lambda t, cu_seqlens, freqs: R._thd(t, cu_seqlens, freqs, False),
        backward_thd=lambda g, cu_seqlens, freqs: R._thd(g, cu_seqlens, freqs, True),
        forward_2d=lambda t, cos_h, sin_h, cos_w, sin_w: R._2d(t, cos_h.shape[1], cos_w.shape[1], cos_h, sin_h, cos_w, sin_w, False),
        backward_2d=lambda g, cos_h, sin_h, cos_w, sin_w: R._2d(g, cos_h.shape[1], cos_w.shape[1], cos_h, sin_h, cos_w, sin_w, True),
    )


def _xentropy_mod():
    from .contrib.xentropy.softmax_xentropy import SoftmaxCrossEntropyLoss as X

    class _Ctx:
        def save_for_backward(self, *t):
            self.saved_tensors = t

    def forward(logits, labels, smoothing, half_to_float):
        ctx = _Ctx()
        losses = X.forward(ctx, logits, labels, smoothing, -1, half_to_float)  # padding is applied by the python layer of the reference
        return losses, ctx.saved_tensors[1]

    def backward(grad_loss, logits, max_log_sum_exp, labels, smoothing):
        ctx = _Ctx()
        ctx.saved_tensors = (logits.contiguous(), max_log_sum_exp, labels.contiguous().view(-1).to(torch.int64))
        ctx.smoothing, ctx.padding_idx = smoothing, -1
        return X.backward(ctx, grad_loss)[0]

    return _mod("xentropy_cuda", forward=forward, backward=backward)


def extension_modules() -> dict:
    from .ops import amp_C
    from .parallel import syncbn_ops
    from .transformer.functional import fused_weight_gradient
    from .utils import flatten

    mods = {"amp_C": amp_C, "syncbn": syncbn_ops, "fused_weight_gradient_mlp_cuda": fused_weight_gradient,
            "apex_C": _mod("apex_C", flatten=flatten.flatten, unflatten=flatten.unflatten),
            "fused_layer_norm_cuda": _layer_norm_mod(), "fused_rotary_positional_embedding": _rope_mod(), "xentropy_cuda": _xentropy_mod()}
    mods.update(_softmax_mods())
    try:
        from .contrib.optimizers import fused_adam_cuda

        mods["fused_adam_cuda"] = fused_adam_cuda
    except Exception:  # noqa: BLE001
        pass
    return mods


def install(overwrite: bool = False) -> list[str]:
    """Register the extension-name modules in ``sys.modules`` (existing entries are kept unless ``overwrite``); returns the names."""
    done = []
    for name, m in extension_modules().items():
        if overwrite or name not in sys.modules:
            sys.modules[name] = m
            done.append(name)
    return done
