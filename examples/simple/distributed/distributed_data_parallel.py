"""Minimal distributed example (reference examples/simple/distributed/distributed_data_parallel.py): one process per GPU, fake data,
apex_b200.parallel.DistributedDataParallel (flat-bucket all-reduce overlapped with backward) + FusedSGD + dynamic loss scaling.

    python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node 2 distributed_data_parallel.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
from apex_b200.optimizers import FusedSGD  # noqa: E402
from apex_b200.parallel import DistributedDataParallel  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", 0))
distributed = int(os.environ.get("WORLD_SIZE", 1)) > 1
use_cuda = torch.cuda.is_available()          # without a GPU the same script runs on gloo / CPU (PyTorch reference paths)
device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
if use_cuda:
    torch.cuda.set_device(local_rank)
if distributed:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.distributed.init_process_group(backend="nccl" if use_cuda else "gloo", init_method="env://")

N, D_in, D_out = 64, 1024, 16
x = torch.randn(N, D_in, device=device)
y = torch.randn(N, D_out, device=device)
model = torch.nn.Linear(D_in, D_out).to(device)
optimizer = FusedSGD(model.parameters(), lr=1e-3, momentum=0.9)
scaler = torch.amp.GradScaler(device.type)
if distributed:
    model = DistributedDataParallel(model)
loss_fn = torch.nn.MSELoss()
for t in range(int(os.environ.get("STEPS", 500))):
    optimizer.zero_grad()
    with torch.autocast(device.type, dtype=torch.float16 if use_cuda else torch.bfloat16):
        loss = loss_fn(model(x).float(), y)
    scaler.scale(loss).backward()
    scaler.step(optimizer)
    scaler.update()
if local_rank == 0:
    print("final loss = ", loss.item())
if distributed:
    torch.distributed.destroy_process_group()
