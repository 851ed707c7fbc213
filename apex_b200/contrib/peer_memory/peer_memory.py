"""``PeerMemoryPool`` on the symmetric heap (parallel/symmetric.py). Reference: apex/contrib/peer_memory/peer_memory.py:6-115 — a
cudaIpc blob per rank with a static + a dynamic bump allocator and typed views into every peer's blob
(peer_memory_cuda.allocate_raw / get_raw_ipc_address / get_raw_peers / blob_view_*)."""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from ...parallel.symmetric import SymmetricMemory, _tensor_from_ptr


class PeerMemoryPool:
    def __init__(self, static_size, dynamic_size, peer_ranks=None, group=None):
        rank = dist.get_rank()
        world_size = dist.get_world_size()
        ngpus = min(torch.cuda.device_count(), world_size)
        base = (rank // ngpus) * ngpus
        if peer_ranks is None:
            peer_ranks = [i + base for i in range(ngpus)]
        for pr in peer_ranks:
            assert base <= pr < base + ngpus, f"{rank} :: peer_rank {pr} not on same node (ranks=[{base},{base + ngpus - 1}])"
        self.alignment = 256
        self.static_size = (static_size + 255) // 256 * 256
        self.dynamic_size = (dynamic_size + 255) // 256 * 256
        self.peer_ranks = peer_ranks
        if group is None and len(peer_ranks) != world_size:
            group = dist.new_group(ranks=peer_ranks)
        self.group = group
        self.mem = SymmetricMemory(self.static_size + self.dynamic_size, group=group, multicast=False, tag="pool")
        self.raw = self.mem.local_ptr
        self.peer_raw = list(self.mem.peer_ptrs)
        self.static_offset = 0
        self.dynamic_offset = 0

    def reset(self):
        self.dynamic_offset = 0

    def allocate_peer_tensors(self, shape, dtype, channels_last, dynamic):
        """One tensor per peer rank (index = rank in the peer group), all at the same offset of the respective blobs."""
        nbytes = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        if dynamic:
            start = (self.dynamic_offset + self.alignment - 1) // self.alignment * self.alignment
            self.dynamic_offset = start + nbytes
            assert self.dynamic_offset < self.dynamic_size, "Dynamic peer memory pool exhausted"
            off = self.static_size + start
        else:
            start = (self.static_offset + self.alignment - 1) // self.alignment * self.alignment
            self.static_offset = start + nbytes
            assert self.static_offset < self.static_size, "Static peer memory pool exhausted"
            off = start
        out = []
        for p in self.peer_raw:
            t = _tensor_from_ptr(p + off, nbytes, self.mem.device, self.mem).view(dtype).view(*shape)
            if channels_last and len(shape) == 4:
                n, c, h, w = shape
                t = t.view(n, h, w, c).permute(0, 3, 1, 2)
            out.append(t)
        return out
