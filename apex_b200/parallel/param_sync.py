"""Registry of parameter buffers whose all-gather may still be in flight (``DistributedFusedAdam(overlap_param_sync=True)``).

The optimizer's step kernels (csrc/dist_adam.cu) release a per-(bucket, rank) flag when a rank has pushed its shard of a bucket into
every rank's parameter buffer. Consumers never wait for "the all-gather": a weight tile is usable as soon as the buckets that hold
its rows have landed.
  * flag-aware GEMMs (``apex_b200.ops.gemm`` — FusedDense / FusedDenseGeluDense / MLP forward): the kernel's TMA producer acquires the
    flags of each B tile right before loading it (``lookup``);
  * everything else: ``wait(tensor)`` enqueues stream-ordered ``cuStreamWaitValue32`` waits on the buckets under that tensor — no
    kernel, no spinning SM (used by the module pre-forward hooks of ``attach_param_sync_hooks`` and by the GEMM front-end for layouts
    it cannot guard tile by tile).
Replaces the reference's pre-forward hook + NCCL all-gather pipeline (apex/contrib/optimizers/distributed_fused_adam.py:938-1071,
:2463-2501)."""
from __future__ import annotations

import torch

from .. import _lib

_lib.declare("ab_stream_wait_geq", "p i p")


class Region:
    """One parameter buffer: ``[base, base + nbytes)`` holds ``bucket_elems``-element buckets of ``esize``-byte elements."""

    def __init__(self, buffer: torch.Tensor, flags_ptr: int, world: int, bucket_elems: int):
        self.base, self.nbytes, self.esize = buffer.data_ptr(), buffer.numel() * buffer.element_size(), buffer.element_size()
        self.flags_ptr, self.world, self.bucket_elems = flags_ptr, world, bucket_elems
        self.epoch = 0            # value the flags reach when the most recent parameter push has landed
        self.in_flight = False
        self.device = buffer.device

    def contains(self, t: torch.Tensor) -> bool:
        p = t.data_ptr()
        return self.base <= p < self.base + self.nbytes

    def buckets_of(self, t: torch.Tensor):
        off = (t.data_ptr() - self.base) // self.esize
        span = (t.untyped_storage().nbytes() // self.esize) if not t.is_contiguous() else t.numel()
        return off, range(off // self.bucket_elems, (off + max(span, 1) - 1) // self.bucket_elems + 1)

    def wait(self, t: torch.Tensor, stream=None) -> None:
        """Work enqueued on ``stream`` (default: current) after this call runs only once every bucket under ``t`` has landed."""
        if not self.in_flight:
            return
        st = _lib.stream_ptr(self.device) if stream is None else stream.cuda_stream
        _, buckets = self.buckets_of(t)
        for b in buckets:
            for r in range(self.world):
                _lib.fn("ab_stream_wait_geq")(self.flags_ptr + (b * 8 + r) * 4, self.epoch & 0x7FFFFFFF, st)


_regions: list[Region] = []


def register(region: Region) -> Region:
    _regions.append(region)
    return region


def unregister(region: Region) -> None:
    if region in _regions:
        _regions.remove(region)


def find(t: torch.Tensor):
    for r in _regions:
        if r.in_flight and r.contains(t):
            return r
    return None


def lookup(w: torch.Tensor):
    """For a K-major weight [N, K] inside an in-flight buffer: (flags_ptr, epoch, world, element offset, bucket_elems); else None."""
    r = find(w)
    if r is None:
        return None
    off, _ = r.buckets_of(w)
    return r, (r.flags_ptr, r.epoch & 0x7FFFFFFF, r.world, off, r.bucket_elems)


def wait(t: torch.Tensor) -> None:
    r = find(t)
    if r is not None:
        r.wait(t)


def attach_param_sync_hooks(model: torch.nn.Module) -> list:
    """Forward-pre hooks for every leaf module that owns parameters and is NOT flag-aware: the current stream waits (stream-ordered,
    per bucket) for exactly the buckets under that module's parameters. Flag-aware modules (FusedDense, FusedDenseGeluDense, MLP)
    synchronise tile by tile inside their GEMM kernels and get no hook."""
    from ..fused_dense import FusedDense, FusedDenseGeluDense
    from ..mlp import MLP

    aware = (FusedDense, FusedDenseGeluDense, MLP)
    handles = []

    def pre(mod, args):
        for p in mod.parameters(recurse=False):
            wait(p)

    for m in model.modules():
        if isinstance(m, aware) or not any(True for _ in m.parameters(recurse=False)):
            continue
        handles.append(m.register_forward_pre_hook(pre))
    return handles
