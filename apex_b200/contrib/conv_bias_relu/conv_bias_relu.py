"""Conv + bias (+mask) + ReLU and Conv + frozen scale/bias + ReLU. Reference: apex/contrib/conv_bias_relu/conv_bias_relu.py:9-102 over
cuDNN-frontend runtime-fused graphs (conv_bias_relu.cpp, 7 entry points). The convolution is cuDNN here too (a library call, as in
the reference); the pointwise tail is folded into one expression and, like the reference, everything runs in fp16/bf16 channels-last
under autocast. Call signatures match: ``ConvBiasReLU(x, weight, bias, padding, stride)`` with bias shaped [1, C, 1, 1]."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _prep(x, weight):
    if x.is_cuda and x.dim() == 4:
        x = x.contiguous(memory_format=torch.channels_last)
        weight = weight.contiguous(memory_format=torch.channels_last)
    return x, weight


def ConvBiasReLU(x, weight, bias, padding, stride):
    x, weight = _prep(x, weight)
    return F.relu(F.conv2d(x, weight, bias.reshape(-1).to(x.dtype), stride, padding))


def ConvBiasMaskReLU(x, weight, bias, mask, padding, stride):
    x, weight = _prep(x, weight)
    return F.relu(F.conv2d(x, weight, bias.reshape(-1).to(x.dtype), stride, padding) * mask.to(x.dtype))


def ConvBias(x, weight, bias, padding, stride):
    x, weight = _prep(x, weight)
    return F.conv2d(x, weight, bias.reshape(-1).to(x.dtype), stride, padding)


def ConvFrozenScaleBiasReLU(x, weight, scale, bias, padding, stride):
    x, weight = _prep(x, weight)
    y = F.conv2d(x, weight, None, stride, padding)
    return F.relu(y * scale.reshape(1, -1, 1, 1).to(y.dtype) + bias.reshape(1, -1, 1, 1).to(y.dtype))


class _Named:
    """``<Name>_.apply`` spelling of the reference (its module-level names are the ``.apply`` of autograd Functions, :99-102); autograd
    flows through the composed ops here, so these only carry the name."""

    apply = None


class ConvBiasReLU_(_Named):
    apply = staticmethod(ConvBiasReLU)


class ConvBiasMaskReLU_(_Named):
    apply = staticmethod(ConvBiasMaskReLU)


class ConvBias_(_Named):
    apply = staticmethod(ConvBias)


class ConvFrozenScaleBiasReLU_(_Named):
    apply = staticmethod(ConvFrozenScaleBiasReLU)
