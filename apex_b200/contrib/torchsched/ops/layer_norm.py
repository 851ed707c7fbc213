"""LayerNorm as graph-level ops: forward returns ``(y, mean, invstd)``, backward takes them back. Reference:
apex/contrib/torchsched/ops/layer_norm.py:268-472 (``cudnn::layer_norm`` / ``cudnn::layer_norm_backward`` custom ops over cached cuDNN
graphs, with fake kernels and a registered autograd formula).

Here the ops are ``apex_b200::norm_fwd`` / ``apex_b200::norm_bwd`` (normalization/custom_ops.py: fake kernels and autograd registered there,
CUDA tensors run csrc/layer_norm_{fwd,bwd}.cu); this module gives them the reference's call signatures. There is no cuDNN handle or graph
cache to manage (the reference's ``CuDNNManager`` / ``LayerNormGraphFactory``): the kernels take the stream from the caller and need no plan."""
from __future__ import annotations

import math

import torch

from ....normalization import custom_ops as _ops

__all__ = ["layer_norm", "layer_norm_fake", "layer_norm_backward", "layer_norm_backward_fake"]


def _check(x, normalized_shape):
    if tuple(x.shape[-len(normalized_shape):]) != tuple(normalized_shape):
        raise ValueError(f"layer_norm expects x.shape[{-len(normalized_shape)}:] == normalized_shape, got {tuple(x.shape)=}, {tuple(normalized_shape)=}")


def layer_norm(x, normalized_shape, weight, bias, eps: float = 1e-5):
    """-> (y, x_mean[rows], x_invstd[rows]); differentiable, traceable by dynamo / AOT autograd as one node."""
    _check(x, normalized_shape)
    return _ops.norm_fwd_op(x.contiguous(), weight, bias, list(normalized_shape), float(eps), False, False)


def layer_norm_fake(x, normalized_shape, weight, bias, eps: float = 1e-5):
    """Shapes and dtypes of :func:`layer_norm`'s results without computing them."""
    rows = math.prod(x.shape[: x.dim() - len(normalized_shape)])
    return torch.empty_like(x), x.new_empty(rows, dtype=torch.float32), x.new_empty(rows, dtype=torch.float32)


def layer_norm_backward(d_y, x_mean, x_invstd, x, normalized_shape, weight, bias):
    """-> (d_x, d_weight, d_bias) from the statistics the forward returned (argument order of the reference, :359-367)."""
    _check(x, normalized_shape)
    return _ops.norm_bwd_op(d_y.contiguous(), x, x_mean, x_invstd, weight, bias, list(normalized_shape), 0.0, False, False, x.dtype)


def layer_norm_backward_fake(d_y, x_mean, x_invstd, x, normalized_shape, weight, bias):
    return torch.empty_like(x), torch.empty_like(weight), torch.empty_like(bias)
