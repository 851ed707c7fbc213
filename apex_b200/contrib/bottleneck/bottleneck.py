"""ResNet bottleneck block with frozen BN folded into per-channel scale/bias, and its spatially (H-) parallel variant.
Reference: apex/contrib/bottleneck/bottleneck.py:32-1410 over ``fast_bottleneck`` (26 cuDNN-frontend fused graphs, 3.6k lines).
Convolutions are cuDNN here as in the reference; every scale / bias / residual / ReLU tail is ONE hand-written kernel in place over
the convolution output, and every backward tail (ReLU mask, residual gradient, scaled gradient for cuDNN's dgrad / wgrad, per-channel
reductions) is one kernel too (contrib/conv_bias_relu -> csrc/conv_epilogue.cu), behind custom autograd Functions.
``SpatialBottleneck`` shards the activation along H over ``spatial_group_size`` ranks: the one-row halo exchange around the 3x3
convolution runs on a side stream WHILE the interior rows are convolved; only the two boundary rows wait for the halos. Its backward
sends the halo-row gradients back to the neighbours the same way, overlapped with the weight gradient (reference
apex/contrib/bottleneck/bottleneck.py:304-830: three side streams, same split)."""
from __future__ import annotations

import functools

import torch
import torch.nn.functional as F  # noqa: F401  (re-exported for users of the reference's module namespace)
from torch import nn

from ..conv_bias_relu.conv_bias_relu import _cl, epilogue_bwd, epilogue_fwd, fused_conv_epilogue


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

    def _conv2(self, out, s2, b2):
        return fused_conv_epilogue(out, self.conv2.weight, bias=b2, scale=s2, stride=1, padding=1, relu=True)

    def forward(self, x):
        x = self._to_nchw(x)
        (s1, b1), (s2, b2), (s3, b3) = self._folded(0, self.bn1), self._folded(1, self.bn2), self._folded(2, self.bn3)
        out = fused_conv_epilogue(x, self.conv1.weight, bias=b1, scale=s1, stride=self.conv1.stride, padding=0, relu=True)
        out = self._conv2(out, s2, b2)
        if self.downsample is None:
            identity = x
        else:
            s4, b4 = (self.w_scale[3], self.w_bias[3]) if self.w_scale is not None else self.downsample[1].get_scale_bias()
            identity = fused_conv_epilogue(x, self.downsample[0].weight, bias=b4, scale=s4, stride=self.downsample[0].stride, padding=0,
                                           relu=False)
        out = fused_conv_epilogue(out, self.conv3.weight, bias=b3, scale=s3, z=identity, stride=1, padding=0, relu=True)
        return self._from_nchw(out)


class _SpatialConv3x3(torch.autograd.Function):
    """relu(conv3x3(x with one-row halos from the H-neighbours) * scale + bias) on an H-shard.

    forward : the halo rows travel on a side stream while cuDNN convolves the rows that do not need them (H padding 0 -> H - 2 output
              rows); the top / bottom output rows are convolved from 3-row slabs once the halos have landed; one fused epilogue kernel.
    backward: fused drelu / dscale kernel; dgrad over the halo-extended input; the two halo rows of that gradient belong to the
              neighbours and are sent back (side stream, overlapped with wgrad), the rows that arrive are added to this shard's edges."""

    @staticmethod
    def forward(ctx, x, weight, scale, bias, halo_ex):
        x, weight = _cl(x), _cl(weight)
        w = weight.to(x.dtype)
        cuda = x.is_cuda
        top_out, btm_out = x[:, :, :1, :].contiguous(), x[:, :, -1:, :].contiguous()
        if cuda:
            cur, side = torch.cuda.current_stream(x.device), halo_ex.stream1
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                top_in, btm_in = halo_ex.left_right_halo_exchange(top_out, btm_out)
        else:
            top_in, btm_in = halo_ex.left_right_halo_exchange(top_out, btm_out)
        H = x.shape[2]
        conv = lambda t: torch.ops.aten.convolution(t, w, None, (1, 1), (0, 1), (1, 1), False, (0, 0), 1)  # noqa: E731
        inner = conv(x) if H > 2 else None                       # output rows 1 .. H-2 need no halo
        if cuda:
            cur.wait_stream(side)
            top_in.record_stream(cur)
            btm_in.record_stream(cur)
        xp = _cl(torch.cat((top_in.to(x.dtype), x, btm_in.to(x.dtype)), dim=2))     # kept for wgrad
        top = conv(xp[:, :, :3, :]) if H > 1 else None
        btm = conv(xp[:, :, -3:, :]) if H > 1 else None
        if H == 1:
            y = conv(xp)
        else:
            y = torch.cat([t for t in (top, inner, btm) if t is not None], dim=2)
        y = _cl(y)
        sc32, b32 = scale.detach().reshape(-1).float().contiguous(), bias.detach().reshape(-1).float().contiguous()
        out = epilogue_fwd(y, sc32, b32, None, None, True, False)
        ctx.save_for_backward(xp, weight, out, sc32)
        ctx.halo_ex = halo_ex
        return out

    @staticmethod
    def backward(ctx, dout):
        xp, weight, out, sc32 = ctx.saved_tensors
        halo_ex = ctx.halo_ex
        need = ctx.needs_input_grad
        dy, _, _, _ = epilogue_bwd(_cl(dout), out, None, None, sc32, True, False, False, False)
        w = weight.to(xp.dtype)
        # dgrad first: its halo rows must leave as early as possible
        dxp, _, _ = torch.ops.aten.convolution_backward(dy, xp, w, None, (1, 1), (0, 1), (1, 1), False, (0, 0), 1, (True, False, False))
        g_top, g_btm = dxp[:, :, :1, :].contiguous(), dxp[:, :, -1:, :].contiguous()
        cuda = dy.is_cuda
        if cuda:
            cur, side = torch.cuda.current_stream(dy.device), halo_ex.stream1
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                r_top, r_btm = halo_ex.left_right_halo_exchange(g_top, g_btm)
        else:
            r_top, r_btm = halo_ex.left_right_halo_exchange(g_top, g_btm)
        dw = None
        if need[1]:
            _, dw, _ = torch.ops.aten.convolution_backward(dy, xp, w, None, (1, 1), (0, 1), (1, 1), False, (0, 0), 1, (False, True, False))
            dw = dw.to(weight.dtype)
        dx = dxp[:, :, 1:-1, :].clone(memory_format=torch.channels_last)
        if cuda:
            cur.wait_stream(side)
            r_top.record_stream(cur)
            r_btm.record_stream(cur)
        dx[:, :, :1, :] += r_top.to(dx.dtype)      # the upper neighbour's gradient for the row it borrowed from this shard
        dx[:, :, -1:, :] += r_btm.to(dx.dtype)
        return dx, dw, None, None, None


class SpatialBottleneck(Bottleneck):
    """Bottleneck whose activations are split along H across ``spatial_group_size`` ranks; a one-row halo exchange feeds the 3x3 conv."""

    def __init__(self, in_channels, bottleneck_channels, out_channels, stride=1, groups=1, dilation=1, norm_func=None, use_cudnn=False,
                 explicit_nhwc=False, spatial_parallel_args=None):
        super().__init__(in_channels, bottleneck_channels, out_channels, stride, groups, dilation, norm_func, use_cudnn, explicit_nhwc)
        self.spatial_parallel_args = spatial_parallel_args  # (spatial_group_size, spatial_group_rank, spatial_communicator, halo_ex, method)
        self.conv2_nopad_h = None

    def _conv2(self, out, s2, b2):
        args = self.spatial_parallel_args
        if args is None or args[0] <= 1:
            return super()._conv2(out, s2, b2)
        return _SpatialConv3x3.apply(out, self.conv2.weight, s2, b2, args[3])


# ---- functional entry points with the reference's argument lists -------------------------------------------------------------------------
def _bottleneck_core(nhwc, stride_1x1, scale, bias, x, conv, conv2):
    """The block as a function of explicit tensors: ``conv`` = (w1, w2, w3[, w4 of the downsample branch]), ``scale`` / ``bias`` the folded
    frozen-BN pairs in the same order; weights are [K, C, R, S] ([K, R, S, C] when ``nhwc``), ``x`` is NCHW-shaped (NHWC when ``nhwc``)."""
    vec = lambda t: t.reshape(1, -1, 1, 1)                                              # noqa: E731
    if nhwc:
        x = x.permute(0, 3, 1, 2)
        conv = [w.permute(0, 3, 1, 2) for w in conv]
    s, b = [vec(t) for t in scale], [vec(t) for t in bias]
    out = fused_conv_epilogue(x, conv[0], bias=b[0], scale=s[0], stride=stride_1x1, padding=0, relu=True)
    out = conv2(out, conv[1], s[1], b[1])
    identity = x
    if len(conv) > 3:
        identity = fused_conv_epilogue(x, conv[3], bias=b[3], scale=s[3], stride=stride_1x1, padding=0, relu=False)
    out = fused_conv_epilogue(out, conv[2], bias=b[2], scale=s[2], z=identity, stride=1, padding=0, relu=True)
    return out.permute(0, 2, 3, 1) if nhwc else out


class BottleneckFunction:
    """``BottleneckFunction.apply(nhwc, stride_1x1, scale, bias, x, *conv)`` (reference bottleneck.py:80-132 over ``fast_bottleneck``): the
    whole block from explicit weights and folded BN vectors. Differentiable through the per-convolution autograd Functions of
    contrib/conv_bias_relu (fused epilogue forward, fused drelu / dscale backward), so no block-level backward is needed."""

    @staticmethod
    def apply(nhwc, stride_1x1, scale, bias, x, *conv):
        plain = lambda t, w, s, b: fused_conv_epilogue(t, w, bias=b, scale=s, stride=1, padding=1, relu=True)     # noqa: E731
        return _bottleneck_core(nhwc, stride_1x1, scale, bias, x, list(conv), plain)


bottleneck_function = BottleneckFunction.apply


class SpatialBottleneckFunction:
    """``SpatialBottleneckFunction.apply(spatial_group_size, spatial_group_rank, spatial_communicator, spatial_halo_exchanger,
    spatial_method, use_delay_kernel, explicit_nhwc, stride_1x1, scale, bias, thresholdTop, thresholdBottom, x, *conv)`` (reference
    bottleneck.py:304-605). ``spatial_method`` / ``use_delay_kernel`` / the thresholds select between the reference's three halo
    strategies; here there is one (halo exchange on a side stream under the interior convolution, see :class:`_SpatialConv3x3`)."""

    @staticmethod
    def apply(spatial_group_size, spatial_group_rank, spatial_communicator, spatial_halo_exchanger, spatial_method, use_delay_kernel,
              explicit_nhwc, stride_1x1, scale, bias, thresholdTop, thresholdBottom, x, *conv):
        if spatial_group_size > 1:
            conv2 = lambda t, w, s, b: _SpatialConv3x3.apply(t, w, s, b, spatial_halo_exchanger)                   # noqa: E731
        else:
            conv2 = lambda t, w, s, b: fused_conv_epilogue(t, w, bias=b, scale=s, stride=1, padding=1, relu=True)  # noqa: E731
        return _bottleneck_core(explicit_nhwc, stride_1x1, scale, bias, x, list(conv), conv2)


spatial_bottleneck_function = SpatialBottleneckFunction.apply
