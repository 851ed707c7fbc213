"""All-reduce a tensor that lives in an NCCL-registered memory pool (reference apex/contrib/examples/nccl_allocator/allreduce.py), and the
B200-native alternative: a tensor on this library's symmetric heap, which its in-kernel collectives use directly.
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 examples/contrib/nccl_allocator/allreduce.py"""
import os

import torch
import torch.distributed as dist

import apex_b200.contrib.nccl_allocator as nccl_allocator


def main():
    nccl_allocator.init()                      # NCCL_NVLS_ENABLE etc. must be set before the communicator exists
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl")
    pool = nccl_allocator.create_nccl_mem_pool()
    with nccl_allocator.nccl_mem(pool):
        a = torch.ones(1 << 20, device="cuda")  # allocated by ncclMemAlloc, registered with the communicator on exit
    dist.all_reduce(a)
    torch.cuda.synchronize()
    assert float(a[0]) == dist.get_world_size()
    sym, mem = nccl_allocator.symmetric_empty(1 << 20, torch.float32)
    sym.fill_(dist.get_rank())
    dist.barrier()
    torch.cuda.synchronize()
    if dist.get_rank() == 0:
        print("all-reduce on an NCCL pool tensor ok; symmetric heap tensor", tuple(sym.shape), "multicast:", mem.has_multicast)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
