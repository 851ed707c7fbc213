"""1-D halo exchange through peer memory. Reference: apex/contrib/peer_memory/peer_halo_exchanger_1d.py:5-84 over
``push_pull_halos_1d`` (peer_memory_cuda.cu:146-295: 16-byte "flit" stores carrying payload + flag into the neighbour's buffer and
a volatile spin on the local one). Here each rank writes its two outgoing halos straight into the neighbours' transfer buffers
with P2P copies over NVLink, a device-side epoch barrier on the signal pad (no NCCL, no host sync) orders them, and the incoming
halos are copied from the local transfer buffers into the padded tensor."""
from __future__ import annotations

import torch

from ...parallel.symmetric import SignalPad


class PeerHaloExchanger1d:
    def __init__(self, ranks, rank_in_group, peer_pool, half_halo):
        self.peer_group_size = len(ranks)
        self.ranks = ranks
        self.peer_rank = rank_in_group
        self.low_neighbor = (self.peer_rank + self.peer_group_size - 1) % self.peer_group_size
        self.high_neighbor = (self.peer_rank + 1) % self.peer_group_size
        self.low_zero = self.peer_rank == 0
        self.high_zero = self.peer_rank == self.peer_group_size - 1
        self.peer_pool = peer_pool
        self.half_halo = half_halo
        self.pad = SignalPad.get(peer_pool.group, peer_pool.mem.device)

    def _slices(self, y, H_split, explicit_nhwc):
        hh = self.half_halo
        dim = (1 if explicit_nhwc else 2) if H_split else (2 if explicit_nhwc else 3)
        L = y.shape[dim] - 2 * hh
        sl = lambda a, b: tuple(slice(a, b) if d == dim else slice(None) for d in range(4))
        return y[sl(hh, 2 * hh)], y[sl(0, hh)], y[sl(L, L + hh)], y[sl(L + hh, L + 2 * hh)]

    def __call__(self, y, H_split=True, explicit_nhwc=False, numSM=0, diagnostics=False):
        low_out, low_in, high_out, high_in = self._slices(y, H_split, explicit_nhwc)
        # tx[r] is the buffer living on rank r: [0] receives from its low neighbour, [1] from its high neighbour
        tx = self.peer_pool.allocate_peer_tensors([2] + list(low_out.shape), y.dtype, False, True)
        if not self.low_zero:
            tx[self.low_neighbor][1].copy_(low_out)    # my low-side interior rows are the low neighbour's HIGH input halo
        if not self.high_zero:
            tx[self.high_neighbor][0].copy_(high_out)
        self.pad.barrier(channel=40)                    # every push has landed everywhere
        mine = tx[self.peer_rank]
        if self.low_zero:
            low_in.zero_()
        else:
            low_in.copy_(mine[0])
        if self.high_zero:
            high_in.zero_()
        else:
            high_in.copy_(mine[1])
        self.pad.barrier(channel=41)                    # buffers may be reused after this
