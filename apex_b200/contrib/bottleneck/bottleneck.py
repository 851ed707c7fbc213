"""ResNet bottleneck block with frozen BN folded into per-channel scale/bias, and its spatially (H-) parallel variant.
Reference: apex/contrib/bottleneck/bottleneck.py:32-1410 over ``fast_bottleneck`` (26 cuDNN-frontend fused graphs, 3.6k lines).
Convolutions are cuDNN here as in the reference; the scale/bias/ReLU/residual tails are single fused pointwise expressions.
``SpatialBottleneck`` shards the activation along H over ``spatial_group_size`` ranks and exchanges one-row halos around the 3x3
convolution through a :mod:`halo_exchangers` transport (peer memory over NVLink by default)."""
from __future__ import annotations

import functools

import torch
import torch.nn.functional as F
from torch import nn


def kaiming_uniform_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    nn.init.kaiming_uniform_(tensor, a=a, mode=mode, nonlinearity=nonlinearity)


def compute_scale_bias_one(nhwc, weight, bias, running_mean, running_var, w_scale, w_bias):
    """Fold one frozen BN into (scale, bias), written INTO the given tensors (reference :20-24): capturable in a CUDA graph."""
    scale = weight * running_var.rsqrt()
    w_scale.copy_(scale)
    w_bias.copy_(bias - running_mean * scale)


def compute_scale_bias_method(nhwc, args):
    for arg in args:
        compute_scale_bias_one(nhwc, *arg)


def drelu_dscale1(grad_o, output, scale1):
    """ReLU backward followed by one scale: (grad * mask * scale1, grad * mask)."""
    dx_relu = (output > 0) * grad_o
    return dx_relu * scale1, dx_relu


def drelu_dscale2(grad_o, output, scale1, scale2):
    dx_relu = (output > 0) * grad_o
    return dx_relu * scale1, dx_relu * scale2


class FrozenBatchNorm2d(nn.Module):
    """BatchNorm2d with fixed statistics and affine parameters, exposed as a per-channel (scale, bias)."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def get_scale_bias(self, nhwc=False):
        scale = self.weight * self.running_var.rsqrt()
        bias = self.bias - self.running_mean * scale
        shape = (1, 1, 1, -1) if nhwc else (1, -1, 1, 1)
        return scale.reshape(shape), bias.reshape(shape)

    def forward(self, x):
        scale, bias = self.get_scale_bias(False)
        return x * scale.to(x.dtype) + bias.to(x.dtype)


def conv3x3(in_planes, out_planes, stride=1, groups=1, dilation=1):
    return nn.Conv2d(in_planes, out_planes, 3, stride, dilation, groups=groups, bias=False, dilation=dilation)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, 1, stride, bias=False)


class Bottleneck(nn.Module):
    """ResNet v1.5-style block with the stride on the first 1x1 convolution (as the reference); frozen BN only."""

    def __init__(self, in_channels, bottleneck_channels, out_channels, stride=1, groups=1, dilation=1, norm_func=None, use_cudnn=False,
                 explicit_nhwc=False):
        super().__init__()
        if groups != 1:
            raise RuntimeError("Only support groups == 1")
        if dilation != 1:
            raise RuntimeError("Only support dilation == 1")
        if norm_func is not None:
            raise RuntimeError("Only support frozen BN now.")
        norm_func = FrozenBatchNorm2d
        self.downsample = None
        if stride != 1 or in_channels != out_channels:
            self.downsample = nn.Sequential(conv1x1(in_channels, out_channels, stride), norm_func(out_channels))
        self.conv1 = conv1x1(in_channels, bottleneck_channels, stride)
        self.conv2 = conv3x3(bottleneck_channels, bottleneck_channels)
        self.conv3 = conv1x1(bottleneck_channels, out_channels)
        self.bn1, self.bn2, self.bn3 = norm_func(bottleneck_channels), norm_func(bottleneck_channels), norm_func(out_channels)
        self.stride, self.use_cudnn, self.explicit_nhwc = stride, use_cudnn, explicit_nhwc
        self.w_scale = self.w_bias = None
        for w in (self.conv1, self.conv2, self.conv3):
            kaiming_uniform_(w.weight, a=1)

    def get_scale_bias_callable(self):
        """Allocates persistent folded (scale, bias) tensors for bn1..bn3 (+ the downsample BN) and returns a callable that refreshes them from
        the BN buffers; forward() then uses the persistent tensors. The reference's hook for recomputing the folding inside a captured graph
        (:233-248)."""
        self.w_scale, self.w_bias, args = [], [], []
        norms = [self.bn1, self.bn2, self.bn3] + ([self.downsample[1]] if self.downsample is not None else [])
        for bn in norms:
            sc = torch.empty_like(bn.weight)
            bi = torch.empty_like(sc)
            args.append((bn.weight, bn.bias, bn.running_mean, bn.running_var, sc, bi))
            self.w_scale.append(sc.reshape(1, -1, 1, 1))
            self.w_bias.append(bi.reshape(1, -1, 1, 1))
        return functools.partial(compute_scale_bias_method, self.explicit_nhwc, args)

    def _folded(self, i, bn):
        return (self.w_scale[i], self.w_bias[i]) if self.w_scale is not None else bn.get_scale_bias()

    def _to_nchw(self, x):
        return x.permute(0, 3, 1, 2) if self.explicit_nhwc else x

    def _from_nchw(self, x):
        return x.permute(0, 2, 3, 1) if self.explicit_nhwc else x

    def _conv2(self, out):
        return self.conv2(out)

    def forward(self, x):
        x = self._to_nchw(x)
        (s1, b1), (s2, b2), (s3, b3) = self._folded(0, self.bn1), self._folded(1, self.bn2), self._folded(2, self.bn3)
        out = F.relu(self.conv1(x) * s1.to(x.dtype) + b1.to(x.dtype))
        out = F.relu(self._conv2(out) * s2.to(x.dtype) + b2.to(x.dtype))
        out = self.conv3(out) * s3.to(x.dtype) + b3.to(x.dtype)
        if self.downsample is None:
            identity = x
        elif self.w_scale is not None:
            identity = self.downsample[0](x) * self.w_scale[3].to(x.dtype) + self.w_bias[3].to(x.dtype)
        else:
            identity = self.downsample(x)
        return self._from_nchw(F.relu(out + identity))


class SpatialBottleneck(Bottleneck):
    """Bottleneck whose activations are split along H across ``spatial_group_size`` ranks; a one-row halo exchange feeds the 3x3 conv."""

    def __init__(self, in_channels, bottleneck_channels, out_channels, stride=1, groups=1, dilation=1, norm_func=None, use_cudnn=False,
                 explicit_nhwc=False, spatial_parallel_args=None):
        super().__init__(in_channels, bottleneck_channels, out_channels, stride, groups, dilation, norm_func, use_cudnn, explicit_nhwc)
        self.spatial_parallel_args = spatial_parallel_args  # (spatial_group_size, spatial_group_rank, spatial_communicator, halo_ex, method)
        self.conv2_nopad_h = None

    def _conv2(self, out):
        args = self.spatial_parallel_args
        if args is None or args[0] <= 1:
            return self.conv2(out)
        halo_ex = args[3]
        top_out, btm_out = out[:, :, :1, :].contiguous(), out[:, :, -1:, :].contiguous()
        top_in, btm_in = halo_ex.left_right_halo_exchange(top_out, btm_out)
        padded = torch.cat((top_in, out, btm_in), dim=2)
        # halos replace the H padding: pad W only
        return F.conv2d(padded, self.conv2.weight, None, self.conv2.stride, (0, 1))
