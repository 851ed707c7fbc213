"""``GroupBatchNorm2d`` — batch norm over groups of ``group_size`` GPUs. Reference: apex/contrib/cudnn_gbn/batch_norm.py:85-216 over a
cuDNN-frontend ``BN_FINALIZE``-with-peers graph and PeerMemoryPool buffers (cudnn_gbn.cpp:26-45). Same module contract on top of the
fused SyncBatchNorm kernel; input is NCHW-shaped, channels-last memory format (what the reference asserts)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...parallel.sync_batchnorm import SyncBatchNorm, create_syncbn_process_group


class GroupBatchNorm2d(SyncBatchNorm):
    def __init__(self, num_features, group_size, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        group = None
        if group_size > 1 and dist.is_initialized() and dist.get_world_size() > group_size:
            group = create_syncbn_process_group(group_size)
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats,
                         process_group=group)
        self.group_size = group_size

    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError(f"expected 4D input (got {input.dim()}D input)")

    def forward(self, input):
        if input.is_cuda and not input.is_contiguous(memory_format=torch.channels_last):
            raise ValueError("GroupBatchNorm2d expects channels_last input")
        return super().forward(input)
