from __future__ import annotations

import torch
import torch.distributed as dist

_groups: dict = {}


def get_unique_nccl_id(n: int = 1):
    """The reference broadcasts an ncclUniqueId to build a second communicator; torch.distributed already owns one."""
    return torch.zeros(n, 128, dtype=torch.uint8)


def init_nccl_comm(unique_id, my_rank: int, num_ranks: int, group=None):
    handle = len(_groups)
    _groups[handle] = group
    return handle


def add_delay(delay: int):
    """Inject artificial latency on the current stream (reference add_delay kernel, nccl_p2p_cuda.cu:19-33)."""
    torch.cuda._sleep(int(delay))


def left_right_halo_exchange_inplace(handle, low_zero, high_zero, low_out_halo, high_out_halo, low_inp_halo, high_inp_halo):
    group = _groups.get(handle)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ops = []
    glob = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    lo, hi = (rank - 1) % world, (rank + 1) % world
    lo_o, hi_o = low_out_halo.contiguous(), high_out_halo.contiguous()
    lo_i, hi_i = torch.empty_like(lo_o), torch.empty_like(hi_o)
    if not low_zero:
        ops += [dist.P2POp(dist.isend, lo_o, glob(lo), group), dist.P2POp(dist.irecv, lo_i, glob(lo), group)]
    if not high_zero:
        ops += [dist.P2POp(dist.isend, hi_o, glob(hi), group), dist.P2POp(dist.irecv, hi_i, glob(hi), group)]
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    low_inp_halo.zero_() if low_zero else low_inp_halo.copy_(lo_i)
    high_inp_halo.zero_() if high_zero else high_inp_halo.copy_(hi_i)


def left_right_halo_exchange(handle, low_zero, high_zero, low_out_halo, high_out_halo):
    low_inp_halo, high_inp_halo = torch.empty_like(low_out_halo), torch.empty_like(high_out_halo)
    left_right_halo_exchange_inplace(handle, low_zero, high_zero, low_out_halo, high_out_halo, low_inp_halo, high_inp_halo)
    return low_inp_halo, high_inp_halo
