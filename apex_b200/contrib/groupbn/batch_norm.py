"""``BatchNorm2d_NHWC`` — NHWC batch norm (+ residual add + ReLU) whose statistics span ``bn_group`` in {1, 2, 4, 8} GPUs of a node.
Reference: apex/contrib/groupbn/batch_norm.py:290-468 over ``bnp`` (4.6k lines of NHWC kernels that exchange partial sums through
cudaIpc buffers with a log2(bn_group)-step butterfly and a magic-number flag, nhwc_batch_norm_kernel.h:358-460).

On B200 this IS the fused SyncBatchNorm kernel (csrc/syncbn.cu): one persistent kernel per direction, statistics pushed into
every peer's exchange buffer over NVLink and merged in one step — an NVSwitch makes the butterfly unnecessary."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...parallel.sync_batchnorm import SyncBatchNorm, SyncBatchnormFunction

# bn_group size -> the process group of this rank with that many members (filled by BatchNorm2d_NHWC); the functional entry points below
# receive only the integer, as the reference's do (there the peers are implied by the IPC pointers)
_bn_groups: dict = {}


def _group_of(bn_group: int):
    if bn_group <= 1:
        return None
    if bn_group not in _bn_groups:
        raise RuntimeError(f"no process group of size bn_group={bn_group} yet: construct a BatchNorm2d_NHWC(bn_group={bn_group}) first "
                           "(it builds the groups collectively)")
    return _bn_groups[bn_group]


def _bn_nhwc(x, z, s, b, rm, riv, mom, epsilon, fuse_relu, is_train, bn_group):
    xv = x.permute(0, 3, 1, 2)
    zv = None if z is None else z.permute(0, 3, 1, 2)
    if is_train:
        y = SyncBatchnormFunction.apply(xv, zv, s, b, rm, riv, epsilon, True, mom, _group_of(bn_group), fuse_relu)
    else:
        y = torch.nn.functional.batch_norm(xv, rm, riv, s, b, False, 0.0, epsilon)
        if zv is not None:
            y = y + zv
        y = torch.relu(y) if fuse_relu else y
    return y.permute(0, 2, 3, 1)


class bn_NHWC_impl:
    """Functional form with the reference's 23-argument list (groupbn/batch_norm.py:8-160): ``apply(x, s, b, rm, riv, mini_m, mini_riv,
    ret_cta, mom, epsilon, fuse_relu, is_train, bn_group, my_data, pair_data, magic, pair_data2, pair_data3, fwd_occup, fwd_grid_x,
    bwd_occup, bwd_grid_x, multi_stream)`` on [N, H, W, C] tensors; ``riv`` is the running variance. Differentiable w.r.t. x, s, b.
    The bnp plumbing (scratch ``mini_*`` / ``ret_cta``, IPC pointers, magic, occupancies, grid sizes) is accepted and unused: the fused
    kernel keeps its statistics in the autograd node, finds its peers through the symmetric heap and sizes its own persistent grid."""

    @staticmethod
    def apply(x, s, b, rm, riv, mini_m=None, mini_riv=None, ret_cta=None, mom=0.1, epsilon=1e-5, fuse_relu=False, is_train=True,
              bn_group=1, my_data=None, pair_data=None, magic=None, pair_data2=None, pair_data3=None, fwd_occup=None, fwd_grid_x=None,
              bwd_occup=None, bwd_grid_x=None, multi_stream=False):
        return _bn_nhwc(x, None, s, b, rm, riv, mom, epsilon, fuse_relu, is_train, bn_group)


class bn_addrelu_NHWC_impl:
    """relu(bn(x) + z) with the reference's 24-argument list (groupbn/batch_norm.py:163-287): ``apply(x, z, s, b, rm, riv, mini_m,
    mini_riv, grid_dim_y, ret_cta, mom, epsilon, is_train, bn_group, my_data, pair_data, magic, pair_data2, pair_data3, fwd_occup,
    fwd_grid_x, bwd_occup, bwd_grid_x, multi_stream)``; gradients flow to x, z, s, b (the ReLU mask is recomputed in the backward kernel
    instead of being stored as the reference's bitmask)."""

    @staticmethod
    def apply(x, z, s, b, rm, riv, mini_m=None, mini_riv=None, grid_dim_y=None, ret_cta=None, mom=0.1, epsilon=1e-5, is_train=True,
              bn_group=1, my_data=None, pair_data=None, magic=None, pair_data2=None, pair_data3=None, fwd_occup=None, fwd_grid_x=None,
              bwd_occup=None, bwd_grid_x=None, multi_stream=False):
        return _bn_nhwc(x, z, s, b, rm, riv, mom, epsilon, True, is_train, bn_group)


class BatchNorm2d_NHWC(SyncBatchNorm):
    def __init__(self, num_features, fuse_relu=False, bn_group=1, max_cta_per_sm=2, cta_launch_margin=12, multi_stream=False):
        group = None
        if bn_group > 1:
            assert dist.is_initialized(), "bn_group > 1 needs torch.distributed"
            world, rank = dist.get_world_size(), dist.get_rank()
            assert world % bn_group == 0
            for g in range(world // bn_group):
                ranks = list(range(g * bn_group, (g + 1) * bn_group))
                pg = dist.new_group(ranks=ranks)
                if rank in ranks:
                    group = pg
        super().__init__(num_features, process_group=group, channel_last=True, fuse_relu=fuse_relu)
        self.bn_group = bn_group
        if group is not None:
            _bn_groups.setdefault(bn_group, group)

    def forward(self, x, z=None):
        """x (and the optional residual z) are [N, H, W, C] tensors; returns relu?(bn(x) + z) in the same layout."""
        if x.dim() != 4:
            raise ValueError("BatchNorm2d_NHWC expects [N, H, W, C] input")
        xv = x.permute(0, 3, 1, 2)  # logical NCHW view of the NHWC storage: exactly the channels-last layout the kernel reads
        zv = None if z is None else z.permute(0, 3, 1, 2)
        return super().forward(xv, zv).permute(0, 2, 3, 1)


def bn_group_is_local(m) -> bool:
    return m.bn_group == 1
