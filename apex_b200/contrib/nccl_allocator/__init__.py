"""``nccl_allocator`` (reference apex/contrib/nccl_allocator/nccl_allocator.py:18-82: a CUDAPluggableAllocator over ncclMemAlloc exposed
as a torch.cuda.MemPool so NCCL can register user buffers / use NVLS zero-copy). In this library collective-facing buffers come from
:class:`apex_b200.parallel.symmetric.SymmetricMemory` (cuMem VMM + NVSwitch multicast), which is what those registrations exist to
enable; the pool API is kept so reference call sites run unchanged."""
from .nccl_allocator import create_nccl_mem_pool, init, nccl_mem, symmetric_empty  # noqa: F401
