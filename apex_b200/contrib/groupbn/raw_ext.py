"""Raw ``bnp`` entry points (reference apex/contrib/csrc/groupbn/interface.cpp:76-99): NHWC batch norm (+ residual add + ReLU) as
separate forward / backward calls that keep their state in caller-owned tensors — ``minibatch_mean`` / ``minibatch_inv_var`` written by
the forward and read by the backward, and for the add-ReLU pair a ``bitmask`` of the ReLU decisions (the backward does not get ``z``).
``BatchNorm2d_NHWC`` here does not go through them (it runs the fused SyncBN kernel, csrc/syncbn.cu); they serve code written against the
extension. Differences by construction: the peers of a ``bn_group`` are a torch.distributed process group registered by
``BatchNorm2d_NHWC(bn_group=...)`` instead of cudaIpc pointer pairs, so the IPC helpers return placeholders and ``my_data`` /
``pair_data*`` / ``magic`` / occupancy / grid arguments are accepted and unused; the bitmask layout is this module's own (one bit per
element in storage order), opaque to the caller like the reference's."""
from __future__ import annotations

import torch
import torch.distributed as dist


def _group(bn_group):
    from .batch_norm import _group_of

    return _group_of(int(bn_group))


def _reduce(t, group):
    if group is not None:
        dist.all_reduce(t, group=group)
    return t


def _stats(x, group):
    """Per-channel mean, biased variance and element count over N, H, W of every rank in the group."""
    xf = x.float()
    C = x.shape[-1]
    packed = torch.cat([xf.sum((0, 1, 2)), (xf * xf).sum((0, 1, 2)), torch.tensor([float(x.numel() // C)], device=x.device)])
    packed = _reduce(packed, group)
    count = packed[-1]
    mean = packed[:C] / count
    var = (packed[C:2 * C] / count - mean * mean).clamp_min_(0.0)
    return mean, var, count


def _pack_bits(mask, out):
    flat = mask.reshape(-1).to(torch.int64)
    pad = (-flat.numel()) % 32
    if pad:
        flat = torch.cat([flat, flat.new_zeros(pad)])
    words = (flat.view(-1, 32) << torch.arange(32, device=flat.device)).sum(1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    assert out.numel() >= words.numel(), "bitmask tensor too small"
    out.view(-1)[:words.numel()].copy_(words)


def _unpack_bits(bitmask, shape):
    n = 1
    for d in shape:
        n *= d
    words = bitmask.view(-1)[:(n + 31) // 32].to(torch.int64)
    bits = (words[:, None] >> torch.arange(32, device=words.device)) & 1
    return bits.reshape(-1)[:n].view(shape).bool()


def _fwd_train(x, z, scale, bias, running_mean, running_var, minibatch_mean, minibatch_inv_var, bitmask, momentum, epsilon, relu,
               bn_group):
    group = _group(bn_group)
    mean, var, count = _stats(x, group)
    inv = torch.rsqrt(var + epsilon)
    minibatch_mean.copy_(mean)
    minibatch_inv_var.copy_(inv)
    running_mean.mul_(1.0 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
    running_var.mul_(1.0 - momentum).add_((var * count / (count - 1).clamp_min(1.0)).to(running_var.dtype), alpha=momentum)
    y = (x.float() - mean) * (inv * scale.float()) + bias.float()
    if z is not None:
        y = y + z.float()
    if relu:
        if bitmask is not None:
            _pack_bits(y > 0, bitmask)
        y = torch.relu(y)
    return y.to(x.dtype)


def _fwd_eval(x, z, scale, bias, running_mean, running_var, epsilon, relu):
    y = (x.float() - running_mean.float()) * (torch.rsqrt(running_var.float() + epsilon) * scale.float()) + bias.float()
    if z is not None:
        y = y + z.float()
    return (torch.relu(y) if relu else y).to(x.dtype)


def _bwd(x, dy, scale, bias, minibatch_mean, minibatch_inv_var, relu_mask, bn_group):
    group = _group(bn_group)
    C = x.shape[-1]
    g = dy.float()
    if relu_mask is not None:
        g = g * relu_mask.to(g.dtype)
    xhat = (x.float() - minibatch_mean) * minibatch_inv_var
    dbias, dscale = g.sum((0, 1, 2)), (g * xhat).sum((0, 1, 2))            # this rank's parameter gradients (the caller's DDP sums them)
    packed = _reduce(torch.cat([dbias, dscale, torch.tensor([float(x.numel() // C)], device=x.device)]), group)
    count = packed[-1]
    dx = (g - packed[:C] / count - xhat * (packed[C:2 * C] / count)) * (minibatch_inv_var * scale.float())
    return dx.to(x.dtype), g.to(x.dtype), dscale.to(scale.dtype), dbias.to(bias.dtype)


def bn_fwd_nhwc(x, scale, bias, running_mean, running_inv_var, minibatch_mean, minibatch_inv_var, ret_cta, momentum, epsilon, fuse_relu,
                my_data, pair_data, pair_data2, pair_data3, bn_group, magic_tensor, occupancy, grid_dim_x, coop):
    return _fwd_train(x, None, scale, bias, running_mean, running_inv_var, minibatch_mean, minibatch_inv_var, None, momentum, epsilon,
                      fuse_relu, bn_group)


def bn_fwd_eval_nhwc(x, scale, bias, running_mean, running_inv_var, ret_cta, bn_group, momentum, epsilon, fuse_relu):
    return _fwd_eval(x, None, scale, bias, running_mean, running_inv_var, epsilon, fuse_relu)


def bn_bwd_nhwc(x, dy, scale, bias, running_mean, running_inv_var, minibatch_mean, minibatch_inv_var, ret_cta, momentum, epsilon,
                fuse_relu, my_data, pair_data, pair_data2, pair_data3, bn_group, magic_tensor, occupancy, grid_dim_x, coop):
    mask = None
    if fuse_relu:   # no bitmask in this signature: the sign of the pre-activation is recomputed from the saved statistics
        mask = ((x.float() - minibatch_mean) * (minibatch_inv_var * scale.float()) + bias.float()) > 0
    dx, _, dscale, dbias = _bwd(x, dy, scale, bias, minibatch_mean, minibatch_inv_var, mask, bn_group)
    return [dx, dscale, dbias]


def bn_addrelu_fwd_nhwc(x, z, scale, bias, running_mean, running_inv_var, minibatch_mean, minibatch_inv_var, bitmask, ret_cta, momentum,
                        epsilon, my_data, pair_data, pair_data2, pair_data3, bn_group, magic_tensor, occupancy, grid_dim_x, coop):
    return _fwd_train(x, z, scale, bias, running_mean, running_inv_var, minibatch_mean, minibatch_inv_var, bitmask, momentum, epsilon,
                      True, bn_group)


def bn_addrelu_fwd_eval_nhwc(x, z, scale, bias, running_mean, running_inv_var, ret_cta, bn_group, momentum, epsilon):
    return _fwd_eval(x, z, scale, bias, running_mean, running_inv_var, epsilon, True)


def bn_addrelu_bwd_nhwc(x, dy, scale, bias, running_mean, running_inv_var, minibatch_mean, minibatch_inv_var, bitmask, ret_cta, momentum,
                        epsilon, my_data, pair_data, pair_data2, pair_data3, bn_group, magic_tensor, occupancy, grid_dim_x, coop):
    dx, dz, dscale, dbias = _bwd(x, dy, scale, bias, minibatch_mean, minibatch_inv_var, _unpack_bits(bitmask, x.shape), bn_group)
    return [dx, dz, dscale, dbias]


# ---- IPC plumbing of the reference (ipc.cu): nothing to exchange here, the group is a process group ---------------------------------
def get_buffer_size(bn_sync_steps):
    return 4


def get_data_ptr(buffer):
    return buffer.data_ptr()


def get_remote_data_ptr(handle, offset):
    return 0


def close_remote_data(handle):
    return None


def bn_fwd_nhwc_occupancy():
    return 2


def bn_bwd_nhwc_occupancy():
    return 2


def bn_addrelu_fwd_nhwc_occupancy():
    return 2


def bn_addrelu_bwd_nhwc_occupancy():
    return 2


ENTRY_POINTS = ("get_buffer_size", "get_data_ptr", "get_remote_data_ptr", "close_remote_data", "bn_fwd_nhwc", "bn_fwd_eval_nhwc",
                "bn_bwd_nhwc", "bn_fwd_nhwc_occupancy", "bn_bwd_nhwc_occupancy", "bn_addrelu_fwd_nhwc", "bn_addrelu_fwd_eval_nhwc",
                "bn_addrelu_bwd_nhwc", "bn_addrelu_fwd_nhwc_occupancy", "bn_addrelu_bwd_nhwc_occupancy")
