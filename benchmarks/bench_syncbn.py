"""BASELINE config #5: SyncBatchNorm inside ResNet-50 (synthetic ImageNet), img/s and per-BN-layer fwd+bwd time.
Launch with torchrun for N > 1. Compares apex_b200.parallel.SyncBatchNorm (in-kernel NVLink exchange) with torch.nn.SyncBatchNorm
(NCCL all_gather / all_reduce) — the reference snapshot has no python SyncBN layer to install."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=15)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impls", default="ours,torch")
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if "MASTER_ADDR" not in os.environ:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from apex_b200.models.resnet import resnet50
    from apex_b200.parallel import SyncBatchNorm as OurSBN

    results = []
    for impl in a.impls.split(","):
        norm = OurSBN if impl == "ours" else torch.nn.SyncBatchNorm
        torch.manual_seed(0)
        model = resnet50(norm_layer=norm).to(dev).to(memory_format=torch.channels_last)
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
        x = torch.randn(a.batch, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
        y = torch.randint(0, 1000, (a.batch,), device=dev)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(ddp(x), y)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()

        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results.append({"impl": impl, "n_gpus": world, "batch_per_gpu": a.batch, "ms_per_step": float(t), "img_per_s": world * a.batch / float(t) * 1e3})
        del model, ddp, opt

    # per-layer micro-benchmark on the BN shapes of ResNet-50 (bf16, channels_last)
    layer = []
    for (C, HW) in [(64, 112), (64, 56), (256, 56), (128, 28), (512, 28), (256, 14), (1024, 14), (512, 7), (2048, 7)]:
        row = {"C": C, "HW": HW}
        for impl in a.impls.split(","):
            bn = (OurSBN if impl == "ours" else torch.nn.SyncBatchNorm)(C).to(dev)
            xx = torch.randn(a.batch, C, HW, HW, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            dyy = torch.randn_like(xx)
            for _ in range(5):
                bn(xx).backward(dyy)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                bn(xx).backward(dyy)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 20 * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            row[impl + "_us"] = float(t)
        layer.append(row)
    if rank == 0:
        out = {"resnet50": results, "per_layer_fwd_bwd_us": layer}
        print(json.dumps(out))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"bench_syncbn_n{world}.json"), "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
