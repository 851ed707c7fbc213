"""ImageNet-style training with apex_b200 (reference examples/imagenet/main_amp.py, 526 lines): ResNet-50, channels-last, mixed
precision, SyncBatchNorm, DistributedDataParallel, fused optimizer, NVTX / profiler window, checkpoint + resume, throughput meter.

    python main_amp.py --steps 50                                            # one GPU, synthetic data
    python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node 8 main_amp.py --sync-bn --steps 50
"""
import argparse
import os
import shutil
import sys
import tempfile
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from apex_b200.models.resnet import resnet50  # noqa: E402
from apex_b200.optimizers import FusedSGD  # noqa: E402
from apex_b200.parallel import DistributedDataParallel, convert_syncbn_model  # noqa: E402
from apex_b200.utils.profiling import ProfilerWindow, nvtx_range  # noqa: E402


def parse():
    p = argparse.ArgumentParser(description="apex_b200 ImageNet example")
    p.add_argument("--data", default="", help="ImageFolder root with train/ (synthetic data when empty)")
    p.add_argument("-b", "--batch-size", type=int, default=64, help="per-GPU batch")
    p.add_argument("--epochs", type=int, default=1)
    p.add_argument("--steps", type=int, default=100, help="iterations per epoch for synthetic data")
    p.add_argument("--lr", type=float, default=0.1, help="for a global batch of 256; scaled linearly")
    p.add_argument("--momentum", type=float, default=0.9)
    p.add_argument("--weight-decay", type=float, default=1e-4)
    p.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    p.add_argument("--sync-bn", action="store_true", help="convert BatchNorm to apex_b200.parallel.SyncBatchNorm")
    p.add_argument("--torch-ddp", action="store_true", help="use torch.nn.parallel.DistributedDataParallel instead of ours")
    p.add_argument("--prof", type=int, default=-1, help="profile 10 iterations starting at this one (cudaProfilerStart/Stop + NVTX)")
    p.add_argument("--print-freq", type=int, default=10)
    p.add_argument("--resume", default="", help="checkpoint to resume from")
    p.add_argument("--save-dir", default=os.path.join(tempfile.gettempdir(), "apex_b200_imagenet"), help="directory for checkpoints (never the source tree)")
    p.add_argument("--checkpoint", default="checkpoint.pth.tar", help="file name inside --save-dir (or an absolute path)")
    return p.parse_args()


class AverageMeter:
    def __init__(self):
        self.val = self.sum = self.count = self.avg = 0.0

    def update(self, v, n=1):
        self.val, self.sum, self.count = v, self.sum + v * n, self.count + n
        self.avg = self.sum / self.count


def reduce_tensor(t, world):
    t = t.detach().clone()
    dist.all_reduce(t)
    return t / world


def synthetic_loader(batch, steps, device):
    x = torch.randn(batch, 3, 224, 224, device=device).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (batch,), device=device)
    for _ in range(steps):
        yield x, y


def folder_loader(root, batch, distributed):
    import torchvision.datasets as D
    import torchvision.transforms as T

    ds = D.ImageFolder(os.path.join(root, "train"), T.Compose([T.RandomResizedCrop(224), T.RandomHorizontalFlip(), T.ToTensor(),
                                                                  T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])]))
    sampler = torch.utils.data.distributed.DistributedSampler(ds) if distributed else None
    return torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=sampler is None, sampler=sampler, num_workers=8, pin_memory=True, drop_last=True)


def main():
    args = parse()
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    distributed = world > 1
    use_cuda = torch.cuda.is_available()      # without a GPU the script still runs (gloo, PyTorch reference paths): a plumbing check
    dev = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(local_rank)
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if use_cuda:
            dist.init_process_group("nccl", init_method="env://", device_id=dev)
        else:
            dist.init_process_group("gloo", init_method="env://")
    torch.backends.cudnn.benchmark = True

    model = resnet50().to(dev).to(memory_format=torch.channels_last)
    if args.sync_bn:
        model = convert_syncbn_model(model, channel_last=True)
    lr = args.lr * args.batch_size * world / 256.0
    optimizer = FusedSGD(model.parameters(), lr=lr, momentum=args.momentum, weight_decay=args.weight_decay)
    amp_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": None}[args.dtype]
    scaler = torch.amp.GradScaler(dev.type, enabled=args.dtype == "fp16")
    start_epoch = 0
    if args.resume and os.path.isfile(args.resume):
        ck = torch.load(args.resume, map_location=dev)
        model.load_state_dict(ck["state_dict"])
        optimizer.load_state_dict(ck["optimizer"])
        start_epoch = ck["epoch"]
        if local_rank == 0:
            print(f"=> resumed from {args.resume} (epoch {start_epoch})")
    if distributed:
        model = (torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank] if use_cuda else None) if args.torch_ddp else DistributedDataParallel(model))

    prof = ProfilerWindow(args.prof, 10)
    for epoch in range(start_epoch, args.epochs):
        loader = folder_loader(args.data, args.batch_size, distributed) if args.data else synthetic_loader(args.batch_size, args.steps, dev)
        batch_time, losses = AverageMeter(), AverageMeter()
        model.train()
        end = time.time()
        for i, (x, y) in enumerate(loader):
            prof.step(i)
            with nvtx_range(f"Body of iteration {i}"):
                x = x.to(dev, non_blocking=True).contiguous(memory_format=torch.channels_last)
                y = y.to(dev, non_blocking=True)
                with nvtx_range("forward"), torch.autocast(dev.type, dtype=amp_dtype, enabled=amp_dtype is not None):
                    loss = F.cross_entropy(model(x), y)
                optimizer.zero_grad(set_to_none=True)
                with nvtx_range("backward"):
                    scaler.scale(loss).backward()
                with nvtx_range("optimizer.step()"):
                    scaler.step(optimizer)
                    scaler.update()
            if i % args.print_freq == 0:
                rl = reduce_tensor(loss, world) if distributed else loss.detach()
                if use_cuda:
                    torch.cuda.synchronize()
                losses.update(float(rl), x.size(0))
                batch_time.update((time.time() - end) / args.print_freq if i else time.time() - end)
                end = time.time()
                if local_rank == 0:
                    print(f"Epoch [{epoch}][{i}]  Time {batch_time.val:.3f} ({batch_time.avg:.3f})  "
                          f"Speed {world * args.batch_size / batch_time.val:.1f} ({world * args.batch_size / batch_time.avg:.1f}) img/s  "
                          f"Loss {losses.val:.4f} ({losses.avg:.4f})")
        if local_rank == 0:
            net = model.module if hasattr(model, "module") else model
            os.makedirs(args.save_dir, exist_ok=True)
            ckpt = args.checkpoint if os.path.isabs(args.checkpoint) else os.path.join(args.save_dir, args.checkpoint)
            torch.save({"epoch": epoch + 1, "arch": "resnet50", "state_dict": net.state_dict(), "optimizer": optimizer.state_dict()}, ckpt)
            shutil.copyfile(ckpt, os.path.join(args.save_dir, "model_latest.pth.tar"))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
