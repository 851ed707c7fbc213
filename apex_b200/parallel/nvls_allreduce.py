"""In-kernel all-reduce through the NVSwitch (csrc/nvls_allreduce.cu): ``multimem.ld_reduce`` + ``multimem.st`` on a symmetric
staging buffer, epoch barriers on the signal pad, at most 32 CTAs. User: DistributedDataParallel's bucket all-reduce
(``fused_collectives=True``); NCCL remains the default there."""
from __future__ import annotations

import ctypes

import torch

from .. import _lib
from .symmetric import SignalPad, SymmetricMemory, node_local

_lib.declare("ab_nvls_allreduce", "p l p i i i i i p f i i p")


def available(group=None) -> bool:
    if not _lib.available():
        return False
    try:
        _lib.fn("ab_nvls_allreduce")
    except (AttributeError, KeyError):
        return False
    return node_local(group)


class NvlsAllReduce:
    """Owns one symmetric (multicast-bound) staging buffer of ``capacity_bytes`` for a process group."""

    def __init__(self, group, device, capacity_bytes: int):
        self.mem = SymmetricMemory(capacity_bytes, group=group, device=device, multicast=True, tag="ar")
        if not self.mem.has_multicast:
            raise RuntimeError("NVLS multicast is not available for this group")
        self.pad = SignalPad.get(group, self.mem.device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.mem.device)
        self.world, self.rank = self.mem.world, self.mem.rank

    def allreduce_(self, flat: torch.Tensor, scale: float = 1.0, ctas: int = 16) -> torch.Tensor:
        """In-place sum over the group of a contiguous 1-D fp32 / fp16 / bf16 tensor (times ``scale``)."""
        esz = flat.element_size()
        unit = (16 // esz) * self.world
        n = (flat.numel() + unit - 1) // unit * unit
        assert n * esz <= self.mem.nbytes, "bucket larger than the staging buffer"
        stage = self.mem.view(flat.dtype, n)
        stage[:flat.numel()].copy_(flat)
        if n > flat.numel():
            stage[flat.numel():].zero_()
        start, end = self.pad.next_epoch(), None
        _lib.fn("ab_nvls_allreduce")(self.mem.mc_ptr, n, ctypes.addressof(self.pad.ptrs), self.rank, self.world, start, 44, 45,
                                     self.ticket.data_ptr(), float(scale), int(ctas), _lib.dt(flat), _lib.stream_ptr(flat.device))
        flat.copy_(stage[:flat.numel()])
        return flat
