"""``nccl_allocator`` — memory pools NCCL can register (user-buffer registration / NVLS zero-copy). Reference:
apex/contrib/nccl_allocator/nccl_allocator.py:18-82 over ``_apex_nccl_allocator`` (NCCLAllocator.cpp:17-38: a CUDAPluggableAllocator
around ncclMemAlloc / ncclMemFree). torch now ships exactly that allocator on the NCCL backend (``backend.mem_allocator``), so the pool
is built from it; :func:`symmetric_empty` is the B200-native alternative used by this library's own in-kernel collectives."""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist

_pool = None


def init() -> None:
    os.environ.setdefault("NCCL_NVLS_ENABLE", "1")
    os.environ.setdefault("TORCH_NCCL_USE_TENSOR_REGISTER_ALLOCATOR_HOOK", "0")


def create_nccl_mem_pool(symmetric=None):
    """A torch.cuda.MemPool backed by the process group's NCCL allocator when this torch build exposes one, else the default pool."""
    global _pool
    try:
        backend = dist.distributed_c10d._get_default_group()._get_backend(torch.device("cuda"))
        _pool = torch.cuda.MemPool(backend.mem_allocator)
    except Exception:  # noqa: BLE001
        _pool = torch.cuda.MemPool()
    return _pool


@contextlib.contextmanager
def nccl_mem(pool=None, enabled=True, device=None, group=None):
    """Allocate inside the pool (and register it with the process group when supported)."""
    if not enabled:
        yield
        return
    pool = pool or _pool or create_nccl_mem_pool()
    with torch.cuda.use_mem_pool(pool):
        yield
    try:
        backend = (group or dist.distributed_c10d._get_default_group())._get_backend(torch.device("cuda"))
        backend.register_mem_pool(pool)
    except Exception:  # noqa: BLE001
        pass


def symmetric_empty(numel, dtype, group=None, device=None):
    """The B200-native alternative: a tensor on the symmetric heap (peer-mapped, multicast-bound) -> (tensor, SymmetricMemory)."""
    from ...parallel.symmetric import SymmetricMemory

    esz = torch.empty((), dtype=dtype).element_size()
    mem = SymmetricMemory(numel * esz, group=group, device=device)
    return mem.view(dtype, numel), mem
