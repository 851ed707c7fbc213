"""OpenFold attention with pair bias and mask. Reference: apex/contrib/openfold_triton/mha.py:20-470 (Triton flash-style kernel with
bias + mask, ``enable()/disable()`` toggle, ``CanSchTriMHA`` shape predicate). The score softmax runs on the scaled softmax kernel."""
from __future__ import annotations

import torch

from ...transformer.functional import scaled_softmax

_enabled = None


def is_enabled():
    return _enabled


def enable() -> None:
    global _enabled
    _enabled = True


def disable() -> None:
    global _enabled
    _enabled = False


def CanSchTriMHA(in_shape, has_bias=True, inf=1e9, training=True):
    return len(in_shape) in (4, 5) and in_shape[-1] in (16, 32, 64, 128)


def _core(q, k, v, mask, bias, inf):
    scores = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    if bias is not None:
        scores = scores + bias
    if mask is not None:
        scores = scores + (mask.to(scores.dtype) - 1.0) * inf
    shp = scores.shape
    p = scaled_softmax(scores.reshape(-1, 1, shp[-2], shp[-1]), 1.0).view(shp)
    return torch.matmul(p.to(v.dtype), v)


def AttnTri(q, k, v, mask=None, bias=None, inf=1e9, is_training=True):
    """q/k/v [*, heads, seq, dim]; mask broadcastable 1 = keep; bias broadcastable additive."""
    return _core(q, k, v, mask, bias, inf)


def AttnBiasJIT(q, k, v, mask, bias, inf=1e9):
    return _core(q, k, v, mask, bias, inf)


def AttnNoBiasJIT(q, k, v, mask, inf=1e9):
    return _core(q, k, v, mask, None, inf)


def schedule_triton_mha(in_shape, fwd=True):
    """Tile schedule the reference's Triton kernel would use for this shape -> (BLOCK_M, BLOCK_N, num_warps, num_stages). Kept for call-site
    parity (mha.py:90-128 of the reference holds a per-shape table); nothing is scheduled with it here."""
    return (64, 64, 4, 2) if fwd else (128, 64, 8, 1)


class FusedAttenionCoreFunc:
    """Name of the reference's autograd Function (sic); ``AttnTri`` is its ``apply``."""

    apply = staticmethod(AttnTri)
