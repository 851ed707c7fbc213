"""``BatchNorm2d_NHWC`` — NHWC batch norm (+ residual add + ReLU) whose statistics span ``bn_group`` in {1, 2, 4, 8} GPUs of a node.
Reference: apex/contrib/groupbn/batch_norm.py:290-468 over ``bnp`` (4.6k lines of NHWC kernels that exchange partial sums through
cudaIpc buffers with a log2(bn_group)-step butterfly and a magic-number flag, nhwc_batch_norm_kernel.h:358-460).

On B200 this IS the fused SyncBatchNorm kernel (csrc/syncbn.cu): one persistent kernel per direction, statistics pushed into
every peer's exchange buffer over NVLink and merged in one step — an NVSwitch makes the butterfly unnecessary."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...parallel.sync_batchnorm import SyncBatchNorm


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

    def forward(self, x, z=None):
        """x (and the optional residual z) are [N, H, W, C] tensors; returns relu?(bn(x) + z) in the same layout."""
        if x.dim() != 4:
            raise ValueError("BatchNorm2d_NHWC expects [N, H, W, C] input")
        xv = x.permute(0, 3, 1, 2)  # logical NCHW view of the NHWC storage: exactly the channels-last layout the kernel reads
        zv = None if z is None else z.permute(0, 3, 1, 2)
        return super().forward(xv, zv).permute(0, 2, 3, 1)


def bn_group_is_local(m) -> bool:
    return m.bn_group == 1
