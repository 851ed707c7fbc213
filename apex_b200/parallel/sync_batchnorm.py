"""SyncBatchNorm with the cross-GPU reduction INSIDE the kernel (csrc/syncbn.cu).

The reference ships only the kernels (csrc/welford.cu, csrc/syncbn.cpp:71-89) — its python layer was removed — so the module
contract is reconstructed from the stale tests (tests/distributed/synced_batchnorm/*): ``SyncBatchNorm(num_features, eps, momentum,
affine, track_running_stats, process_group=None, channel_last=False, fuse_relu=False)``, ``forward(input, z=None)``, 1-D/2-D/N-D
inputs, uneven per-rank batch sizes, gradients equal to ``nn.BatchNorm`` on the concatenated batch;
``convert_syncbn_model``; ``create_syncbn_process_group``.

Data path on B200: one persistent kernel per direction = local Welford partials -> grid barrier -> every rank stores its
(mean, M2, count) [or (sum dy, sum dy (x-mean), count)] straight into every peer's exchange buffer over NVLink, flips an epoch
flag, merges the D contributions in rank order -> grid barrier -> normalise(+residual add)(+ReLU). No NCCL call, no
all_gather / all_reduce launches (the reference protocol was all_gather(3C floats) + all_reduce(2C floats) around 4 kernels).
Process groups that are not one NVSwitch domain (or CPU tensors) use the plain torch.distributed implementation below.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from .. import _lib
from ..utils import config

_lib.declare("ab_syncbn", "i i p p p p p i i i i p p p p p p p f f p p p p p l p p i p p i i i i p i i i i p")

_BN_CHANNEL = 32  # signal-pad channel (and device epoch counter) owned by SyncBN
_XCHG_C = 4096  # channels the exchange buffer is sized for (grown on demand)


class _GroupState:
    """Per process-group resources for the fused path: signal pad, double-buffered exchange memory, scratch."""

    _states: dict = {}

    def __init__(self, group, device, channels):
        from .symmetric import SignalPad, SymmetricMemory, node_local

        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device
        self.fused = device.type == "cuda" and _lib.available() and (self.world == 1 or (self.world <= 8 and node_local(group)))
        self.cap = max(channels, _XCHG_C)
        self.pad = None
        self.xchg = None
        if self.fused and self.world > 1:
            self.pad = SignalPad.get(group, device)
            self.region = self.world * self.cap * 3
            self.xchg = SymmetricMemory(2 * self.region * 4, group=group, device=device, multicast=False, tag="bn")
            self.xchg_ptrs = self.xchg.peer_ptr_array()
        self.grid_bar = torch.zeros(2, dtype=torch.int32, device=device)
        # [merged 3C][unit counters C][split partials]; zero-initialised (the kernel leaves the counters at zero)
        self.scratch = torch.zeros(4 * self.cap + 24 * self.cap + 3 * 65536, dtype=torch.float32, device=device)
        self.count = torch.zeros(1, dtype=torch.float32, device=device)

    @classmethod
    def get(cls, group, device, channels):
        key = (id(group) if group is not None else 0, str(device))
        st = cls._states.get(key)
        if st is None or st.cap < channels:
            st = cls._states[key] = _GroupState(group, device, channels)
        return st


def _layout(x: torch.Tensor):
    """-> (x_contig, N, C, HW, nhwc)"""
    if x.dim() < 2:
        raise ValueError("SyncBatchNorm expects at least 2-D input (N, C, ...)")
    N, C = x.shape[0], x.shape[1]
    HW = x.numel() // (N * C) if N * C > 0 else 0
    if x.is_contiguous():
        return x, N, C, HW, 0
    if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last):
        return x, N, C, HW, 1
    if x.dim() == 5 and x.is_contiguous(memory_format=torch.channels_last_3d):
        return x, N, C, HW, 1
    return x.contiguous(), N, C, HW, 0


def _call(st: _GroupState, is_bwd, phases, x, dy, z, out, dz, N, C, HW, nhwc, weight, bias, mean, invstd, var_biased, rmean, rvar,
          momentum, eps, grad_w, grad_b, sum_dy, sum_dy_xmu, fuse_relu, exchange: bool):
    world = st.world if (exchange and st.world > 1) else 1
    if world > 1:
        # the epoch (and with it the half of the exchange buffer in use) lives in device memory: the kernel advances it, so a captured
        # CUDA graph signals / waits on a fresh value at every replay
        epoch, epoch_ctr = 0, st.pad.dev_epochs[_BN_CHANNEL:_BN_CHANNEL + 1].data_ptr()
        pads, xchg = ctypes.addressof(st.pad.ptrs), ctypes.addressof(st.xchg_ptrs)
        off, region = 0, st.region
        rank = st.rank
    else:
        epoch, epoch_ctr, pads, xchg, off, region, rank = 0, None, None, None, 0, 0, 0
    _lib.fn("ab_syncbn")(int(is_bwd), int(phases), x.data_ptr(), _lib.ptr(dy), _lib.ptr(z), _lib.ptr(out), _lib.ptr(dz), N, C, HW, nhwc,
                         _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(var_biased), _lib.ptr(rmean),
                         _lib.ptr(rvar), float(momentum), float(eps), _lib.ptr(grad_w), _lib.ptr(grad_b), _lib.ptr(sum_dy),
                         _lib.ptr(sum_dy_xmu), st.scratch.data_ptr(), st.scratch.numel(), st.count.data_ptr(), st.grid_bar.data_ptr(), int(fuse_relu), pads,
                         xchg, off, rank, world, epoch, epoch_ctr, region, config.syncbn_sm_margin(), _BN_CHANNEL, _lib.dt(x), _lib.stream_ptr(x.device))


class SyncBatchnormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, z, weight, bias, running_mean, running_var, eps, track_running_stats, momentum, process_group, fuse_relu):
        x, N, C, HW, nhwc = _layout(input)
        st = _GroupState.get(process_group, x.device, C)
        if not st.fused:
            return _fallback_forward(ctx, x, z, weight, bias, running_mean, running_var, eps, track_running_stats, momentum,
                                     process_group, fuse_relu)
        zz = None
        if z is not None:
            zz = z.contiguous(memory_format=torch.channels_last) if nhwc and z.dim() == 4 else z.contiguous()
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        w32 = weight.float() if weight is not None else None
        b32 = bias.float() if bias is not None else None
        upd = track_running_stats and running_mean is not None
        if upd and running_mean.dtype != torch.float32:
            raise RuntimeError("running statistics must be float32")
        _call(st, 0, 7, x, None, zz, y, None, N, C, HW, nhwc, w32, b32, mean, invstd, None, running_mean if upd else None,
              running_var if upd else None, momentum, eps, None, None, None, None, fuse_relu, True)
        ctx.save_for_backward(x, zz, w32, b32, mean, invstd)
        ctx.meta = (N, C, HW, nhwc, process_group, fuse_relu, eps, weight is not None, bias is not None, z is not None,
                    weight.dtype if weight is not None else None)
        ctx.fused = True
        return y

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.fused:
            return _fallback_backward(ctx, grad_output)
        x, zz, w32, b32, mean, invstd = ctx.saved_tensors
        N, C, HW, nhwc, group, fuse_relu, eps, has_w, has_b, has_z, wdtype = ctx.meta
        st = _GroupState.get(group, x.device, C)
        dy = grad_output.contiguous(memory_format=torch.channels_last) if (nhwc and grad_output.dim() == 4) else grad_output.contiguous()
        dx = torch.empty_like(x)
        dz = torch.empty_like(x) if has_z else None
        gw = torch.empty(C, dtype=torch.float32, device=x.device)
        gb = torch.empty(C, dtype=torch.float32, device=x.device)
        sdy = torch.empty(C, dtype=torch.float32, device=x.device)
        sdx = torch.empty(C, dtype=torch.float32, device=x.device)
        _call(st, 1, 7, x, dy, zz, dx, dz, N, C, HW, nhwc, w32, b32, mean, invstd, None, None, None, 0.0, eps, gw, gb, sdy, sdx,
              fuse_relu, True)
        return (dx, dz, gw.to(wdtype) if has_w else None, gb.to(wdtype) if has_b else None, None, None, None, None, None, None, None)


# ---- plain torch.distributed implementation (multi-node groups / CPU): same math, library collectives -----------------------
def _fallback_forward(ctx, x, z, weight, bias, running_mean, running_var, eps, track, momentum, group, fuse_relu):
    dims = [0] + list(range(2, x.dim()))
    xf = x.float()
    C = x.shape[1]
    n_local = torch.tensor([x.numel() / C], device=x.device, dtype=torch.float32)
    mean_l = xf.mean(dims)
    var_l = xf.var(dims, unbiased=False)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        packed = torch.cat([mean_l, var_l, n_local])
        allp = [torch.empty_like(packed) for _ in range(world)]
        dist.all_gather(allp, packed, group=group)
        means = torch.stack([p[:C] for p in allp])
        vars_ = torch.stack([p[C:2 * C] for p in allp])
        ns = torch.stack([p[2 * C] for p in allp]).view(-1, 1)
        n_tot = ns.sum()
        mean = (means * ns).sum(0) / n_tot
        var = ((vars_ + (means - mean) ** 2) * ns).sum(0) / n_tot
    else:
        n_tot, mean, var = n_local.sum(), mean_l, var_l
    invstd = torch.rsqrt(var + eps)
    if track and running_mean is not None:
        running_mean.mul_(1 - momentum).add_(momentum * mean)
        running_var.mul_(1 - momentum).add_(momentum * var * n_tot / torch.clamp(n_tot - 1, min=1))
    shape = [1, C] + [1] * (x.dim() - 2)
    y = (xf - mean.view(shape)) * invstd.view(shape)
    if weight is not None:
        y = y * weight.float().view(shape)
    if bias is not None:
        y = y + bias.float().view(shape)
    if z is not None:
        y = y + z.float()
    if fuse_relu:
        y = torch.relu(y)
    ctx.save_for_backward(x, z, weight, bias, mean, invstd, y if fuse_relu else None, n_tot.reshape(1))
    ctx.meta = (group, fuse_relu, z is not None)
    ctx.fused = False
    return y.to(x.dtype)


def _fallback_backward(ctx, grad_output):
    x, z, weight, bias, mean, invstd, y, n_tot = ctx.saved_tensors
    group, fuse_relu, has_z = ctx.meta
    C = x.shape[1]
    dims = [0] + list(range(2, x.dim()))
    shape = [1, C] + [1] * (x.dim() - 2)
    g = grad_output.float()
    if fuse_relu:
        g = g * (y > 0)
    xmu = x.float() - mean.view(shape)
    sum_dy = g.sum(dims)
    sum_dy_xmu = (g * xmu).sum(dims)
    gw = sum_dy_xmu * invstd if weight is not None else None
    gb = sum_dy.clone() if bias is not None else None
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        packed = torch.cat([sum_dy, sum_dy_xmu])
        dist.all_reduce(packed, group=group)
        sum_dy, sum_dy_xmu = packed[:C], packed[C:]
    w = weight.float().view(shape) if weight is not None else 1.0
    dx = (g - (sum_dy / n_tot).view(shape) - xmu * (invstd ** 2 * sum_dy_xmu / n_tot).view(shape)) * w * invstd.view(shape)
    return (dx.to(x.dtype), g.to(x.dtype) if has_z else None, gw.to(weight.dtype) if weight is not None else None,
            gb.to(bias.dtype) if bias is not None else None, None, None, None, None, None, None, None)


class SyncBatchNorm(torch.nn.modules.batchnorm._BatchNorm):
    """Batch norm whose statistics span every rank of ``process_group``. In eval mode (or world size 1 on CPU) it is ordinary BN."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, process_group=None,
                 channel_last=False, fuse_relu=False):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats)
        self.process_group = process_group
        self.channel_last = channel_last
        self.fuse_relu = fuse_relu

    def _check_input_dim(self, input):
        if input.dim() < 2:
            raise ValueError(f"expected at least 2D input (got {input.dim()}D input)")

    def _specify_process_group(self, process_group):
        self.process_group = process_group

    def _specify_channel_last(self, channel_last):
        self.channel_last = channel_last

    def forward(self, input, z=None):
        self._check_input_dim(input)
        if self.channel_last and input.dim() == 4 and input.is_contiguous() and not input.is_contiguous(memory_format=torch.channels_last):
            # the reference's channel_last=True means the tensor is physically [N, H, W, C]
            input = input.permute(0, 3, 1, 2)
            if z is not None:
                z = z.permute(0, 3, 1, 2)
            out = self._forward(input, z)
            return out.permute(0, 2, 3, 1)
        return self._forward(input, z)

    def _forward(self, input, z):
        use_batch_stats = self.training or not self.track_running_stats
        if not use_batch_stats:  # eval: ordinary batch norm with the running statistics
            y = torch.nn.functional.batch_norm(input, self.running_mean, self.running_var, self.weight, self.bias, False, 0.0, self.eps)
            if z is not None:
                y = y + z
            return torch.relu(y) if self.fuse_relu else y
        momentum = self.momentum if self.momentum is not None else 0.1
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if self.momentum is None:
                momentum = 1.0 / float(self.num_batches_tracked)
        return SyncBatchnormFunction.apply(input, z, self.weight, self.bias, self.running_mean, self.running_var, self.eps,
                                           self.track_running_stats, momentum, self.process_group, self.fuse_relu)


def convert_syncbn_model(module, process_group=None, channel_last=False):
    """Recursively replace every ``torch.nn.modules.batchnorm._BatchNorm`` by :class:`SyncBatchNorm` (parameters and buffers are
    shared, not copied)."""
    mod = module
    if isinstance(module, torch.nn.modules.instancenorm._InstanceNorm):
        return module
    if isinstance(module, torch.nn.modules.batchnorm._BatchNorm) and not isinstance(module, SyncBatchNorm):
        mod = SyncBatchNorm(module.num_features, module.eps, module.momentum, module.affine, module.track_running_stats, process_group,
                            channel_last=channel_last)
        mod.running_mean = module.running_mean
        mod.running_var = module.running_var
        mod.num_batches_tracked = module.num_batches_tracked
        if module.affine:
            mod.weight = module.weight
            mod.bias = module.bias
        mod.training = module.training
    for name, child in module.named_children():
        mod.add_module(name, convert_syncbn_model(child, process_group=process_group, channel_last=channel_last))
    del module
    return mod


def create_syncbn_process_group(group_size):
    """Split the world into consecutive groups of ``group_size`` ranks and return the one this rank belongs to."""
    if group_size == 0:
        return None
    world_size = dist.get_world_size()
    assert world_size >= group_size and world_size % group_size == 0
    group = None
    for g in range(world_size // group_size):
        ranks = list(range(g * group_size, (g + 1) * group_size))
        cur = dist.new_group(ranks=ranks)
        if dist.get_rank() // group_size == g:
            group = cur
    assert group is not None
    return group
