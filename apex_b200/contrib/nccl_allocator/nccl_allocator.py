"""``nccl_allocator`` — memory pools NCCL can register (user-buffer registration / NVLS zero-copy). Reference:
apex/contrib/nccl_allocator/nccl_allocator.py:18-82 over ``_apex_nccl_allocator`` (NCCLAllocator.cpp:17-38: a CUDAPluggableAllocator
around ncclMemAlloc / ncclMemFree).

Two pools:
  * :func:`create_nccl_mem_pool` — torch's own ncclMemAlloc allocator (``backend.mem_allocator``) for buffers NCCL should register;
  * :func:`create_symmetric_mem_pool` — a ``CUDAPluggableAllocator`` implemented by THIS library (csrc/symm_heap.cpp
    ``ab_symm_pool_malloc`` / ``ab_symm_pool_free``): every tensor allocated inside it lives in a shareable cuMem VMM allocation, and
    :func:`peer_map` turns any such tensor into per-rank peer pointers after the fact (fd exchange + import), which is what this
    library's in-kernel collectives consume. :func:`symmetric_empty` is the explicit one-buffer form."""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist

_pool = None


def init() -> None:
    os.environ.setdefault("NCCL_NVLS_ENABLE", "1")
    os.environ.setdefault("TORCH_NCCL_USE_TENSOR_REGISTER_ALLOCATOR_HOOK", "0")


def get_func_args(func) -> list:
    """Parameter names of ``func`` (reference nccl_allocator.py:11-15); [] for callables without an introspectable signature."""
    import inspect

    try:
        return [p.name for p in inspect.signature(func).parameters.values()]
    except (TypeError, ValueError):
        return []


def _symmetric_kwargs(symmetric, pool_cls) -> dict:
    """``symmetric`` (None = the pool's default) as the keyword this torch build's MemPool understands: upstream calls it ``symmetric``,
    NVIDIA's builds ``symm_mem`` (reference nccl_allocator.py:18-33); neither -> ValueError."""
    if symmetric is None:
        return {}
    names = get_func_args(pool_cls.__init__) + get_func_args(pool_cls)
    for key in ("symmetric", "symm_mem"):
        if key in names:
            return {key: symmetric}
    raise ValueError("symmetric setting with torch.cuda.MemPool requires higher PyTorch version")


def create_nccl_mem_pool(symmetric=None):
    """A torch.cuda.MemPool backed by the process group's NCCL allocator when this torch build exposes one, else the default pool.
    ``symmetric=True / False`` asks for (or forbids) symmetric registration when the pool type has such a switch."""
    global _pool
    kw = _symmetric_kwargs(symmetric, torch.cuda.MemPool)
    try:
        backend = dist.distributed_c10d._get_default_group()._get_backend(torch.device("cuda"))
        _pool = torch.cuda.MemPool(backend.mem_allocator, **kw)
    except Exception:  # noqa: BLE001
        _pool = torch.cuda.MemPool(**kw)
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


_symm_alloc = None
_symm_pool = None


def create_symmetric_mem_pool():
    """torch.cuda.MemPool over the library's shareable-VMM pluggable allocator."""
    global _symm_alloc, _symm_pool
    from ... import _lib

    if not _lib.available():
        raise _lib.gpu_required_error("create_symmetric_mem_pool")
    if _symm_pool is None:
        so = str(_lib._PKG / "_kernels.so")
        _symm_alloc = torch.cuda.memory.CUDAPluggableAllocator(so, "ab_symm_pool_malloc", "ab_symm_pool_free")
        _symm_pool = torch.cuda.MemPool(_symm_alloc.allocator())
    return _symm_pool


@contextlib.contextmanager
def symmetric_mem(pool=None):
    """Allocate inside the symmetric pool: ``with symmetric_mem(): buf = torch.empty(...)``."""
    pool = pool or create_symmetric_mem_pool()
    with torch.cuda.use_mem_pool(pool):
        yield pool


def peer_map(t: torch.Tensor, group=None):
    """Peer pointers of a tensor that was allocated inside the symmetric pool: list (one entry per rank of ``group``) of device addresses
    valid in THIS process that alias the tensor at the same offset of every rank's corresponding allocation. Collective over the group;
    every rank passes its own tensor of the same size allocated in the same order."""
    import ctypes

    from ... import _lib
    from ...parallel import symmetric as S

    _lib.declare("ab_symm_pool_export", "p p p p")
    base, nbytes, fd = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_int(-1)
    _lib.fn("ab_symm_pool_export")(t.data_ptr(), ctypes.addressof(base), ctypes.addressof(nbytes), ctypes.addressof(fd))
    off = t.data_ptr() - int(base.value)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    g = ctypes.c_uint64(0)
    _lib.fn("ab_symm_granularity")(dev, 1, ctypes.addressof(g))
    sizes = S._all_gather_obj(int(nbytes.value), group)
    fds = S._exchange_fds(fd.value, group, "pm")
    ptrs = []
    for r, f in enumerate(fds):
        if r == rank:
            ptrs.append(t.data_ptr())
            continue
        hh, pp = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _lib.fn("ab_symm_import")(dev, f, sizes[r], int(g.value), ctypes.addressof(hh), ctypes.addressof(pp))
        ptrs.append(int(pp.value) + off)
        os.close(f)
    dist.barrier(group=group)
    os.close(fd.value)
    return ptrs


def symmetric_empty(numel, dtype, group=None, device=None):
    """The B200-native alternative: a tensor on the symmetric heap (peer-mapped, multicast-bound) -> (tensor, SymmetricMemory)."""
    from ...parallel.symmetric import SymmetricMemory

    esz = torch.empty((), dtype=dtype).element_size()
    mem = SymmetricMemory(numel * esz, group=group, device=device)
    return mem.view(dtype, numel), mem
