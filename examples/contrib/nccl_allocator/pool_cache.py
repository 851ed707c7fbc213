"""The NCCL memory pool next to the ordinary caching allocator: pool blocks come from ``ncclMemAlloc``, are cached by the pool (a second
round of allocations reuses them) and are released independently of ``torch.cuda.empty_cache()`` of the default allocator. Counterpart of
the reference's apex/contrib/examples/nccl_allocator/{change_cuda_allocator,cache}.py (which read ``nvidia-smi``; this one reads the
driver's free-memory counter).

    python examples/contrib/nccl_allocator/pool_cache.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
import apex_b200.contrib.nccl_allocator as nccl_allocator  # noqa: E402

MB = 1 << 20


def used_mb() -> float:
    free, total = torch.cuda.mem_get_info()
    return (total - free) / MB


def report(label: str, base: float) -> None:
    torch.cuda.synchronize()
    print(f"{label:<44} {used_mb() - base:9.0f} MB in use")


def main():
    if not torch.cuda.is_available():
        print("needs a GPU")
        return
    nccl_allocator.init()
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")                      # context + first allocator segment
    base = used_mb()
    nrep, n_pool, n_plain = 6, 100 * MB // 4, 50 * MB // 4
    pool = nccl_allocator.create_nccl_mem_pool()
    with nccl_allocator.nccl_mem(pool):
        held = [torch.empty(n_pool, device="cuda") for _ in range(nrep)]      # 6 x 100 MB from ncclMemAlloc
    report("after pool allocations (+600)", base)
    plain = [torch.empty(n_plain, device="cuda") for _ in range(nrep)]        # 6 x 50 MB from the default allocator
    report("after default allocations (+300)", base)
    del plain
    torch.cuda.empty_cache()
    report("default allocator emptied (-300)", base)
    del held
    with nccl_allocator.nccl_mem(pool):
        held = [torch.empty(n_pool, device="cuda") for _ in range(nrep)]
    report("pool allocations again: cached blocks (same)", base)
    del held, pool
    torch.cuda.empty_cache()
    report("pool released (-600)", base)


if __name__ == "__main__":
    main()
