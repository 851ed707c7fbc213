"""A small DDP training loop whose parameters and gradients are allocated inside the NCCL memory pool, so NCCL can use zero-copy /
NVLS algorithms on the gradient buckets (reference apex/contrib/examples/nccl_allocator/toy_ddp.py).
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 examples/contrib/nccl_allocator/toy_ddp.py"""
import os

import torch
import torch.distributed as dist

import apex_b200.contrib.nccl_allocator as nccl_allocator
from apex_b200.optimizers import FusedAdam
from apex_b200.parallel import DistributedDataParallel


def main():
    nccl_allocator.init()
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl")
    torch.manual_seed(0)
    pool = nccl_allocator.create_nccl_mem_pool()
    with nccl_allocator.nccl_mem(pool):          # parameters, their gradients and DDP's flat buckets come from the pool
        model = torch.nn.Sequential(torch.nn.Linear(1024, 4096), torch.nn.GELU(), torch.nn.Linear(4096, 1024)).cuda()
        ddp = DistributedDataParallel(model)
        opt = FusedAdam(model.parameters(), lr=1e-3)
        for step in range(10):
            x = torch.randn(64, 1024, device="cuda")
            loss = ddp(x).square().mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            if dist.get_rank() == 0:
                print(f"step {step} loss {loss.item():.4f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
