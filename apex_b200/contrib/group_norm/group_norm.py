"""NHWC GroupNorm with optional fused SiLU/swish (csrc/group_norm.cu).

Reference: apex/contrib/group_norm/group_norm.py:162-457 (``GroupNorm(num_groups, num_channels, eps, affine, act)``; custom ops
``apex::group_norm_nhwc_fprop/bprop``; v1 one/two-pass kernels for 23 channel counts, v2 Blackwell kernels for 14 (HW, C) shapes,
else ``torch_group_norm``). Here one persistent kernel per direction covers every shape; inputs must be channels-last
(logical NCHW, physical NHWC) like the reference; anything else goes through ``torch_group_norm``."""
from __future__ import annotations

import functools

import torch
import torch.nn.functional as F
from torch import nn

from ... import _lib

_lib.declare("ab_group_norm", "i p p p p p i p p p p p l p i i i i i f i i p")

_lib.declare("ab_group_norm_small", "i p p p p p i p p p p p p i i i i f i i p")
_lib.declare("ab_group_norm_stream", "i p p p p p i p p p p p i i i i f i i p")

_MAX_N = 8192  # per-image arrival counters / ready flags live in a fixed control buffer

_state: dict = {}


def torch_group_norm(x, g, w, b, eps, act=""):
    xdtype, wdtype = x.dtype, w.dtype
    if xdtype != wdtype:
        x = x.to(dtype=wdtype)
    y = F.group_norm(x, g, w, b, eps)
    if act in ("silu", "swish"):
        y = F.silu(y)
    if xdtype != wdtype and y.dtype != xdtype:
        y = y.to(dtype=xdtype)
    return y


def _scratch(device, need):
    st = _state.get(device)
    if st is None or st[0].numel() < need:
        st = _state[device] = [torch.empty(max(need, 1 << 20), dtype=torch.float32, device=device),
                               torch.zeros(1 + 2 * _MAX_N, dtype=torch.int32, device=device) if st is None else st[1], 0 if st is None else st[2]]
    return st


def _native_ok(x, w):
    return (x.is_cuda and _lib.available() and x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16, torch.float32)
            and x.is_contiguous(memory_format=torch.channels_last) and (w is None or w.dtype in (x.dtype, torch.float32))
            and x.shape[0] <= _MAX_N)


_small_ok_cache: dict = {}


def _small_ok(is_bwd, HW, C, G, dtype):
    """csrc/group_norm_small.cu: a CTA (or a cluster of 2 / 4 CTAs) owns a whole (image, group) slab when it fits in shared memory."""
    key = (is_bwd, HW, C, G, dtype)
    r = _small_ok_cache.get(key)
    if r is None:
        import ctypes

        f = _lib.raw_fn("ab_group_norm_small_ok", ctypes.c_int, [ctypes.c_int] * 5)
        r = _small_ok_cache[key] = bool(f(int(is_bwd), HW, C, G, _lib.dt(dtype)))
    return r


def _use_stream(is_bwd: bool, x, G: int) -> bool:
    """Two-pass streaming kernels (csrc/group_norm_stream.cu) instead of the slab-per-group kernels. Measured on B200
    (profiles/results/bench_group_norm_thr{0,100000}.json): they only win in the BACKWARD pass once an (image, group) slab exceeds
    ~300 KB (the slab kernels then need clusters of 8 latency-bound CTAs); the forward never wins. APEX_B200_GN_STREAM_MIN_MB (total
    activation MB; 0 = always, large = never) overrides the rule for A/B runs."""
    import os

    if G > 64:
        return False
    v = os.environ.get("APEX_B200_GN_STREAM_MIN_MB")
    if v is not None:
        return x.numel() * x.element_size() >= float(v) * 1e6
    N, C, H, W = x.shape
    return bool(is_bwd) and H * W * (C // G) * x.element_size() >= 300 * 1024


def _native_bwd(x, dy, weight, bias, mean, rstd, G, eps, silu):
    """-> (dx, dgamma, dbeta). The slab / persistent kernels' last CTA writes dgamma / dbeta in the parameters' own dtype (no cast kernels:
    the small shapes are launch-bound); the streaming kernels accumulate with fp32 atomics and are cast afterwards."""
    dx = torch.empty_like(x)
    C = x.shape[1]
    fp32_out = _use_stream(True, x, G) or weight.dtype != bias.dtype
    dt = torch.float32 if fp32_out else weight.dtype
    dg = torch.empty(C, dtype=dt, device=x.device)
    db = torch.empty(C, dtype=dt, device=x.device)
    _launch(True, x, dy, dx, weight, bias, mean, rstd, dg, db, G, eps, silu)
    if fp32_out:
        dg, db = dg.to(weight.dtype), db.to(bias.dtype)
    return dx, dg, db


def _launch(is_bwd, x, dy, out, w, b, mean, rstd, dg, db, G, eps, silu):
    N, C, H, W = x.shape
    if _use_stream(is_bwd, x, G):
        st = _scratch(x.device, N * C * 2 + 64)
        w_fp32 = int(w is not None and w.dtype == torch.float32 and x.dtype != torch.float32)
        _lib.fn("ab_group_norm_stream")(int(is_bwd), x.data_ptr(), _lib.ptr(dy), out.data_ptr(), _lib.ptr(w), _lib.ptr(b), w_fp32, mean.data_ptr(),
                                        rstd.data_ptr(), _lib.ptr(dg), _lib.ptr(db), st[0].data_ptr(), N, H * W, C, G, float(eps), int(silu),
                                        _lib.dt(x), _lib.stream_ptr(x.device))
        return
    if _small_ok(is_bwd, H * W, C, G, x.dtype):
        st = _scratch(x.device, N * C * 2 + 64)
        w_fp32 = int(w is not None and w.dtype == torch.float32 and x.dtype != torch.float32)
        _lib.fn("ab_group_norm_small")(int(is_bwd), x.data_ptr(), _lib.ptr(dy), out.data_ptr(), _lib.ptr(w), _lib.ptr(b), w_fp32, mean.data_ptr(),
                                       rstd.data_ptr(), _lib.ptr(dg), _lib.ptr(db), st[0].data_ptr(), st[1].data_ptr(), N, H * W, C, G, float(eps),
                                       int(silu), _lib.dt(x), _lib.stream_ptr(x.device))
        return
    need = N * C * 16 * 3 + N * C * 3 + N * G * 2 + 64
    st = _scratch(x.device, need)
    scratch, ctrl = st[0], st[1]
    st[2] += 1  # launch epoch: the kernel publishes "image n is normalisable" by writing this value into the image's ready flag
    w_fp32 = int(w is not None and w.dtype == torch.float32 and x.dtype != torch.float32)
    _lib.fn("ab_group_norm")(int(is_bwd), x.data_ptr(), _lib.ptr(dy), out.data_ptr(), _lib.ptr(w), _lib.ptr(b), w_fp32, mean.data_ptr(),
                             rstd.data_ptr(), _lib.ptr(dg), _lib.ptr(db), scratch.data_ptr(), scratch.numel(), ctrl.data_ptr(), st[2] & 0x7FFFFFFF, N, H * W, C, G,
                             float(eps), int(silu), _lib.dt(x), _lib.stream_ptr(x.device))


class GroupNormNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, G, weight, bias, eps, act):
        silu = act in ("silu", "swish")
        N, C = x.shape[0], x.shape[1]
        y = torch.empty_like(x)  # preserves channels_last
        mean = torch.empty(N * G, dtype=torch.float32, device=x.device)
        rstd = torch.empty(N * G, dtype=torch.float32, device=x.device)
        _launch(False, x, None, y, weight, bias, mean, rstd, None, None, G, eps, silu)
        ctx.save_for_backward(x, weight, bias, mean, rstd)
        ctx.G, ctx.eps, ctx.silu = G, eps, silu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous(memory_format=torch.channels_last)
        dx, dg, db = _native_bwd(x, dy, weight, bias, mean, rstd, ctx.G, ctx.eps, ctx.silu)
        return dx, None, dg, db, None, None


def group_norm_nhwc(x, G, weight, bias, eps=1e-5, act=""):
    if not _native_ok(x, weight) or weight is None or bias is None:
        return torch_group_norm(x, G, weight, bias, eps, act)
    return GroupNormNHWC.apply(x, G, weight, bias, eps, act)


class GroupNorm(nn.Module):
    """``torch.nn.GroupNorm`` signature + ``act`` ('' | 'silu' | 'swish'); optimised for channels-last input."""

    __constants__ = ["num_groups", "num_channels", "eps", "affine", "act"]

    def __init__(self, num_groups, num_channels, eps=1e-5, affine=True, device=None, dtype=None, act=None):
        super().__init__()
        if num_channels % num_groups != 0:
            raise ValueError("num_channels must be divisible by num_groups")
        self.num_groups, self.num_channels, self.eps, self.affine = num_groups, num_channels, eps, affine
        self.act = (act or "").lower()
        kw = {"device": device, "dtype": dtype}
        if affine:
            self.weight = nn.Parameter(torch.empty(num_channels, **kw))
            self.bias = nn.Parameter(torch.empty(num_channels, **kw))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.affine:
            nn.init.ones_(self.weight)
            nn.init.zeros_(self.bias)

    def forward(self, input):
        if self.affine and _compiled_cuda(input) and input.dim() == 4 and input.is_contiguous(memory_format=torch.channels_last):
            return group_norm_nhwc_fprop_op(input, self.num_groups, self.weight, self.bias, self.eps, self.act)[0]
        if self.affine and _native_ok(input, self.weight) and not torch.compiler.is_compiling():
            return GroupNormNHWC.apply(input, self.num_groups, self.weight, self.bias, self.eps, self.act)
        return torch_group_norm(input, self.num_groups, self.weight, self.bias, self.eps, self.act)

    def extra_repr(self):
        return "{num_groups}, {num_channels}, eps={eps}, affine={affine}, act={act}".format(**self.__dict__)


# functional entry points of the reference module (group_norm.py:193-245): one implementation serves all three names
def cuda_group_norm_nhwc_one_pass(x, G, weight, bias, eps, act=None):
    return group_norm_nhwc(x, G, weight, bias, eps, act or "")


cuda_group_norm_nhwc_two_pass = cuda_group_norm_nhwc_one_pass
cuda_group_norm_v2_nhwc = cuda_group_norm_nhwc_one_pass


def get_cc_and_sm_count(device_index: int):
    p = torch.cuda.get_device_properties(device_index)
    return (p.major, p.minor), p.multi_processor_count


# fprop / bprop pair of the reference (its torch.library custom ops ``apex::group_norm_nhwc_fprop / _bprop``, group_norm.py:49-190): ``sums``
# is the opaque statistics tensor handed from one to the other (here [2, N * G]: mean and 1 / std per (image, group)).
def group_norm_nhwc_fprop(x, G, weight, bias, eps, act=None, passes=1, use_group_norm_v2=False):
    act = act.lower() if act else ""
    assert x.shape[1] % G == 0, "C % G != 0."
    assert act in ("", "silu", "swish"), "Unsupported activation."
    assert weight.numel() == x.shape[1] and bias.numel() == x.shape[1], "Unexpected parameter count."
    N = x.shape[0]
    sums = torch.empty(2, N * G, dtype=torch.float32, device=x.device)
    if _native_ok(x, weight):
        y = torch.empty_like(x)
        _launch(False, x, None, y, weight, bias, sums[0], sums[1], None, None, G, eps, act in ("silu", "swish"))
        return y, sums
    xf = x.float().reshape(N, G, -1)
    sums[0] = xf.mean(-1).reshape(-1)
    sums[1] = torch.rsqrt(xf.var(-1, unbiased=False) + eps).reshape(-1)
    return torch_group_norm(x, G, weight, bias, eps, act), sums


def group_norm_nhwc_bprop(grad_output, sums, x, G, weight, bias, eps, act=None, passes=1, use_group_norm_v2=False):
    act = act.lower() if act else ""
    if _native_ok(x, weight):
        dy = grad_output.contiguous(memory_format=torch.channels_last)
        return _native_bwd(x, dy, weight, bias, sums[0], sums[1], G, eps, act in ("silu", "swish"))
    # explicit formulas (no autograd: this also runs underneath the autograd dispatch key, as the body of the custom op below)
    N, C = x.shape[0], x.shape[1]
    xf = x.float().reshape(N, G, C // G, -1)
    mean, rstd = sums[0].view(N, G, 1, 1), sums[1].view(N, G, 1, 1)
    xhat = (xf - mean) * rstd
    w, b = weight.float().view(1, G, C // G, 1), bias.float().view(1, G, C // G, 1)
    dz = grad_output.float().reshape(N, G, C // G, -1)
    if act in ("silu", "swish"):
        z = xhat * w + b
        sg = torch.sigmoid(z)
        dz = dz * sg * (1 + z * (1 - sg))
    dxhat = dz * w
    m1 = dxhat.mean((2, 3), keepdim=True)
    m2 = (dxhat * xhat).mean((2, 3), keepdim=True)
    dx = ((dxhat - m1 - xhat * m2) * rstd).reshape(x.shape).to(x.dtype)
    dw = (dz * xhat).sum((0, 3)).reshape(C).to(weight.dtype)
    db = dz.sum((0, 3)).reshape(C).to(bias.dtype)
    return dx, dw, db


# ---- torch.library registration (the reference's ``apex::group_norm_nhwc_fprop / _bprop``, group_norm.py:49-190): torch.compile keeps the fused
# kernel as one graph node instead of tracing the PyTorch fallback.
if hasattr(torch.library, "custom_op"):
    @torch.library.custom_op("apex_b200::group_norm_nhwc_fprop", mutates_args=())
    def group_norm_nhwc_fprop_op(x: torch.Tensor, G: int, weight: torch.Tensor, bias: torch.Tensor, eps: float,
                                 act: str) -> tuple[torch.Tensor, torch.Tensor]:
        y, sums = group_norm_nhwc_fprop(x, G, weight, bias, eps, act)
        return y, sums

    @group_norm_nhwc_fprop_op.register_fake
    def fake_group_norm_nhwc_fprop(x, G, weight, bias, eps, act):
        return torch.empty_like(x), x.new_empty(2, x.shape[0] * G, dtype=torch.float32)

    @torch.library.custom_op("apex_b200::group_norm_nhwc_bprop", mutates_args=())
    def group_norm_nhwc_bprop_op(grad_output: torch.Tensor, sums: torch.Tensor, x: torch.Tensor, G: int, weight: torch.Tensor, bias: torch.Tensor,
                                 eps: float, act: str) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        dx, dw, db = group_norm_nhwc_bprop(grad_output, sums, x, G, weight, bias, eps, act)
        return dx.contiguous(memory_format=torch.channels_last) if dx.dim() == 4 else dx, dw.clone(), db.clone()

    @group_norm_nhwc_bprop_op.register_fake
    def fake_group_norm_nhwc_bprop(grad_output, sums, x, G, weight, bias, eps, act):
        return torch.empty_like(x), torch.empty_like(weight), torch.empty_like(bias)

    def setup_context(ctx, inputs, output):
        x, G, weight, bias, eps, act = inputs
        ctx.save_for_backward(x, weight, bias, output[1])
        ctx.cfg = (G, eps, act)

    def backward(ctx, gy, gsums):
        x, weight, bias, sums = ctx.saved_tensors
        G, eps, act = ctx.cfg
        dx, dw, db = group_norm_nhwc_bprop_op(gy, sums, x, G, weight, bias, eps, act)
        return dx, None, dw, db, None, None

    group_norm_nhwc_fprop_op.register_autograd(backward, setup_context=setup_context)


@functools.cache
def one_time_warning(msg: str) -> None:
    """Warn once per distinct message (reference group_norm.py:22-25)."""
    import warnings

    warnings.warn(msg, stacklevel=2)


def _compiled_cuda(x) -> bool:
    return torch.compiler.is_compiling() and x.is_cuda and hasattr(torch.library, "custom_op")
