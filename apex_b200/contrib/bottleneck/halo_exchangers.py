"""Halo exchangers for spatial (H-split) parallelism. Reference: apex/contrib/bottleneck/halo_exchangers.py:10-276 —
NoComm (perf stub), AllGather, SendRecv (nccl_p2p), Peer (peer_memory flit kernel) and HaloPadder. Every exchanger maps
``left_right_halo_exchange(left_output_halo, right_output_halo[, left_input_halo, right_input_halo])`` to its transport;
the Peer variant writes over NVLink into the neighbours' symmetric buffers and orders with a device-side epoch barrier."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..nccl_p2p import nccl_p2p as _p2p


class HaloExchanger:
    def __init__(self, ranks, rank_in_group):
        self.stream1, self.stream2, self.stream3 = (torch.cuda.Stream() if torch.cuda.is_available() else None for _ in range(3))
        self.group_size = len(ranks)
        self.ranks = ranks
        self.rank_in_group = rank_in_group
        self.wrap_around_left_rank_in_group = (rank_in_group + self.group_size - 1) % self.group_size
        self.wrap_around_right_rank_in_group = (rank_in_group + 1) % self.group_size
        self.left_rank = ranks[rank_in_group - 1] if rank_in_group > 0 else -1
        self.left_zero = rank_in_group == 0
        self.right_rank = ranks[rank_in_group + 1] if rank_in_group < self.group_size - 1 else -1
        self.right_zero = rank_in_group == self.group_size - 1


class HaloExchangerNoComm(HaloExchanger):
    def left_right_halo_exchange(self, left_output_halo, right_output_halo, left_input_halo=None, right_input_halo=None):
        if left_input_halo is None:
            return right_output_halo, left_output_halo
        left_input_halo.copy_(right_output_halo)
        right_input_halo.copy_(left_output_halo)


class HaloExchangerAllGather(HaloExchanger):
    def __init__(self, ranks, rank_in_group, comm):
        super().__init__(ranks, rank_in_group)
        self.comm = comm

    def left_right_halo_exchange(self, left_output_halo, right_output_halo, left_input_halo=None, right_input_halo=None):
        send = torch.stack((left_output_halo, right_output_halo)).contiguous()
        allh = [torch.empty_like(send) for _ in range(self.group_size)]
        dist.all_gather(allh, send, group=self.comm)
        ag_left = allh[self.wrap_around_left_rank_in_group][1]    # left neighbour's right output
        ag_right = allh[self.wrap_around_right_rank_in_group][0]  # right neighbour's left output
        if self.left_zero:
            ag_left = torch.zeros_like(ag_left)
        if self.right_zero:
            ag_right = torch.zeros_like(ag_right)
        if left_input_halo is None:
            return ag_left, ag_right
        left_input_halo.copy_(ag_left)
        right_input_halo.copy_(ag_right)


class HaloExchangerSendRecv(HaloExchanger):
    def __init__(self, ranks, rank_in_group, group=None):
        super().__init__(ranks, rank_in_group)
        self.handle = _p2p.init_nccl_comm(None, rank_in_group, len(ranks), group)

    def left_right_halo_exchange(self, left_output_halo, right_output_halo, left_input_halo=None, right_input_halo=None):
        if left_input_halo is None:
            return _p2p.left_right_halo_exchange(self.handle, self.left_zero, self.right_zero, left_output_halo, right_output_halo)
        _p2p.left_right_halo_exchange_inplace(self.handle, self.left_zero, self.right_zero, left_output_halo, right_output_halo,
                                              left_input_halo, right_input_halo)


class HaloExchangerPeer(HaloExchanger):
    def __init__(self, ranks, rank_in_group, peer_pool, explicit_nhwc, numSM=0):
        super().__init__(ranks, rank_in_group)
        from ...parallel.symmetric import SignalPad

        self.peer_pool, self.explicit_nhwc = peer_pool, explicit_nhwc
        self.pad = SignalPad.get(peer_pool.group, peer_pool.mem.device)

    def left_right_halo_exchange(self, left_output_halo, right_output_halo, left_input_halo=None, right_input_halo=None):
        inplace = left_input_halo is not None
        if not inplace:
            left_input_halo, right_input_halo = torch.empty_like(right_output_halo), torch.empty_like(left_output_halo)
        tx = self.peer_pool.allocate_peer_tensors([2] + list(left_output_halo.shape), left_output_halo.dtype, False, True)
        if not self.left_zero:
            tx[self.wrap_around_left_rank_in_group][1].copy_(left_output_halo)
        if not self.right_zero:
            tx[self.wrap_around_right_rank_in_group][0].copy_(right_output_halo)
        self.pad.barrier(channel=46)
        mine = tx[self.rank_in_group]
        left_input_halo.zero_() if self.left_zero else left_input_halo.copy_(mine[0])
        right_input_halo.zero_() if self.right_zero else right_input_halo.copy_(mine[1])
        self.pad.barrier(channel=47)
        if not inplace:
            return left_input_halo, right_input_halo


class HaloPadder:
    """Pads a tensor along H (or W) with halos received from the neighbours (reference :203-276)."""

    def __init__(self, halo_ex):
        self.halo_ex = halo_ex

    def __call__(self, y, half_halo, explicit_nhwc, H_split):
        dim = (1 if explicit_nhwc else 2) if H_split else (2 if explicit_nhwc else 3)
        L = y.shape[dim]
        shape = list(y.shape)
        shape[dim] = L + 2 * half_halo
        ypad = torch.empty(shape, dtype=y.dtype, device=y.device,
                           memory_format=torch.channels_last if (y.dim() == 4 and y.is_contiguous(memory_format=torch.channels_last)) else torch.contiguous_format)
        sl = lambda a, b: tuple(slice(a, b) if d == dim else slice(None) for d in range(y.dim()))
        ypad[sl(half_halo, half_halo + L)].copy_(y)
        left_out, right_out = y[sl(0, half_halo)].contiguous(), y[sl(L - half_halo, L)].contiguous()
        left_in, right_in = self.halo_ex.left_right_halo_exchange(left_out, right_out)
        ypad[sl(0, half_halo)].copy_(left_in)
        ypad[sl(half_halo + L, L + 2 * half_halo)].copy_(right_in)
        return ypad

    def wait(self):
        pass
