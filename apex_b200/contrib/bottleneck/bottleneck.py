"""ResNet bottleneck block with frozen BN folded into per-channel scale/bias, and its spatially (H-) parallel variant.
Reference: apex/contrib/bottleneck/bottleneck.py:32-1410 over ``fast_bottleneck`` (26 cuDNN-frontend fused graphs, 3.6k lines).
Convolutions are cuDNN here as in the reference; the scale/bias/ReLU/residual tails are single fused pointwise expressions.
``SpatialBottleneck`` shards the activation along H over ``spatial_group_size`` ranks and exchanges one-row halos around the 3x3
convolution through a :mod:`halo_exchangers` transport (peer memory over NVLink by default)."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


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
        for w in (self.conv1, self.conv2, self.conv3):
            nn.init.kaiming_uniform_(w.weight, a=1)

    def _to_nchw(self, x):
        return x.permute(0, 3, 1, 2) if self.explicit_nhwc else x

    def _from_nchw(self, x):
        return x.permute(0, 2, 3, 1) if self.explicit_nhwc else x

    def _conv2(self, out):
        return self.conv2(out)

    def forward(self, x):
        x = self._to_nchw(x)
        s1, b1 = self.bn1.get_scale_bias()
        s2, b2 = self.bn2.get_scale_bias()
        s3, b3 = self.bn3.get_scale_bias()
        out = F.relu(self.conv1(x) * s1.to(x.dtype) + b1.to(x.dtype))
        out = F.relu(self._conv2(out) * s2.to(x.dtype) + b2.to(x.dtype))
        out = self.conv3(out) * s3.to(x.dtype) + b3.to(x.dtype)
        identity = x if self.downsample is None else self.downsample(x)
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
