"""``nccl_p2p_cuda`` — neighbour halo exchange over a communicator of its own. Reference: apex/contrib/csrc/nccl_p2p/nccl_p2p_cuda.cu
:34-128 (``get_unique_nccl_id`` -> broadcast -> ``init_nccl_comm`` -> ``left_right_halo_exchange[_inplace]`` as one grouped
ncclSend / ncclRecv per side on the current stream; ``add_delay`` test kernel).

On CUDA this is the same thing natively (csrc/nccl_p2p.cpp): a real ncclUniqueId, a dedicated ncclComm per handle (so the exchanges
never serialise behind the training job's collectives on torch's communicator), one grouped send/recv pair per neighbour enqueued on
the current stream, no host synchronisation. On CPU tensors (gloo test configuration) the exchange runs through
``torch.distributed.batch_isend_irecv``. The peer-memory exchangers (contrib/peer_memory, NVLink P2P stores) remain the faster path
on one NVSwitch node; this one also works across nodes."""
from __future__ import annotations

import ctypes
import glob
import os

import torch
import torch.distributed as dist

from ... import _lib

_lib.declare("ab_nccl_load", "p")
_lib.declare("ab_nccl_unique_id", "p")
_lib.declare("ab_nccl_comm_init", "p i i p")
_lib.declare("ab_nccl_comm_destroy", "i")
_lib.declare("ab_nccl_exchange", "i i i p p p p l p")

_groups: dict = {}      # handle -> (process group, native comm index or None, rank, world)
_loaded = None


def _native() -> bool:
    """Resolve NCCL once: the copy torch already mapped, else the wheel's library."""
    global _loaded
    if _loaded is None:
        _loaded = False
        if _lib.available():
            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))
            path = (cands[0] if cands else "").encode()
            try:
                _lib.fn("ab_nccl_load")(ctypes.c_char_p(path))
                _loaded = True
            except RuntimeError:
                _loaded = False
    return _loaded


def get_unique_nccl_id(n: int = 1) -> torch.Tensor:
    """[n, 128] uint8: fresh ncclUniqueIds (meaningful on the rank that will broadcast them, as in the reference)."""
    out = torch.zeros(n, 128, dtype=torch.uint8)
    if _native():
        for i in range(n):
            _lib.fn("ab_nccl_unique_id")(out[i].data_ptr())
    return out


def init_nccl_comm(unique_id, my_rank: int, num_ranks: int, group=None) -> int:
    """Build the communicator. ``unique_id``: the [128] / [1, 128] uint8 tensor every rank received from the broadcasting rank, or None —
    then rank 0 of ``group`` creates one and it is broadcast here."""
    handle = len(_groups)
    comm = None
    if _native() and dist.is_initialized() and dist.get_backend(group) == "nccl" and num_ranks > 1:
        if unique_id is None:
            uid = get_unique_nccl_id(1) if my_rank == 0 else torch.zeros(1, 128, dtype=torch.uint8)
            uid = uid.cuda()
            dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            unique_id = uid
        uid = unique_id.detach().reshape(-1)[:128].to("cpu", torch.uint8).contiguous()
        h = ctypes.c_int(-1)
        torch.cuda.synchronize()
        _lib.fn("ab_nccl_comm_init")(uid.data_ptr(), int(my_rank), int(num_ranks), ctypes.addressof(h))
        comm = int(h.value)
    _groups[handle] = (group, comm, my_rank, num_ranks)
    return handle


def destroy_nccl_comm(handle: int) -> None:
    group, comm, _, _ = _groups.pop(handle, (None, None, 0, 0))
    if comm is not None:
        _lib.fn("ab_nccl_comm_destroy")(comm)


def add_delay(delay: int):
    """Inject artificial latency on the current stream (reference add_delay kernel, nccl_p2p_cuda.cu:19-33)."""
    torch.cuda._sleep(int(delay))


def left_right_halo_exchange_inplace(handle, low_zero, high_zero, low_out_halo, high_out_halo, low_inp_halo, high_inp_halo):
    group, comm, my_rank, world = _groups.get(handle, (None, None, None, None))
    if comm is not None and low_out_halo.is_cuda:
        lo_o, hi_o = low_out_halo.contiguous(), high_out_halo.contiguous()
        lo_i = low_inp_halo if low_inp_halo.is_contiguous() else torch.empty_like(lo_o)
        hi_i = high_inp_halo if high_inp_halo.is_contiguous() else torch.empty_like(hi_o)
        nbytes = lo_o.numel() * lo_o.element_size()
        _lib.fn("ab_nccl_exchange")(comm, -1 if low_zero else (my_rank - 1) % world, -1 if high_zero else (my_rank + 1) % world,
                                    lo_o.data_ptr(), lo_i.data_ptr(), hi_o.data_ptr(), hi_i.data_ptr(), nbytes, _lib.stream_ptr(lo_o.device))
        if low_zero:
            low_inp_halo.zero_()
        elif lo_i is not low_inp_halo:
            low_inp_halo.copy_(lo_i)
        if high_zero:
            high_inp_halo.zero_()
        elif hi_i is not high_inp_halo:
            high_inp_halo.copy_(hi_i)
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ops = []
    glob_rank = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    lo, hi = (rank - 1) % world, (rank + 1) % world
    lo_o, hi_o = low_out_halo.contiguous(), high_out_halo.contiguous()
    lo_i, hi_i = torch.empty_like(lo_o), torch.empty_like(hi_o)
    if not low_zero:
        ops += [dist.P2POp(dist.isend, lo_o, glob_rank(lo), group), dist.P2POp(dist.irecv, lo_i, glob_rank(lo), group)]
    if not high_zero:
        ops += [dist.P2POp(dist.isend, hi_o, glob_rank(hi), group), dist.P2POp(dist.irecv, hi_i, glob_rank(hi), group)]
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    low_inp_halo.zero_() if low_zero else low_inp_halo.copy_(lo_i)
    high_inp_halo.zero_() if high_zero else high_inp_halo.copy_(hi_i)


def left_right_halo_exchange(handle, low_zero, high_zero, low_out_halo, high_out_halo):
    low_inp_halo, high_inp_halo = torch.empty_like(low_out_halo), torch.empty_like(high_out_halo)
    left_right_halo_exchange_inplace(handle, low_zero, high_zero, low_out_halo, high_out_halo, low_inp_halo, high_inp_halo)
    return low_inp_halo, high_inp_halo
