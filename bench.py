#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): DistributedFusedAdam step time over the Llama-3-8B parameter set.

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--config llama3-8b]

One process per GPU (the driver launches N>1 through torch.distributed.run). A "step" is a full ZeRO-2 optimizer step on
synthetic gradients that already sit in the contiguous gradient buffer: gradient reduce-scatter over the N ranks + Adam on the
fp32 (param, exp_avg, exp_avg_sq) shard + all-gather of the new bf16 parameters. fp32 state, bf16 grads, bf16 params
(the configuration BASELINE.md §2 row 4 names). Strong scaling: the parameter set is fixed, each rank owns 1/N of the state.

Timing: W untimed warm-up steps, then exactly K steps between two CUDA events on the launching stream, bracketed by
barrier + synchronize; value = max over ranks of ms/step. The working set (>= 44 GB per GPU at N=8, 128 GB at N=1) is far
larger than the 126 MB L2, so no L2 flush is needed between iterations. Clocks are sampled with nvidia-smi during the timed
region. `e2e` repeats the measurement through the public API with, every step, a host->device copy of that step's learning
rate from pinned memory and a device->host read of one updated parameter (the step's result).

--impl reference runs the UNMODIFIED reference (baseline/_ref, NCCL + its own kernels) on the same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="llama3-8b")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fused", default="auto", choices=["auto", "off"], help="ours: in-kernel collectives (auto) or the NCCL path (off)")
    return ap.parse_args()


def main():
    args = parse()
    if args.impl == "reference":
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        if not os.path.isdir(os.path.join(ref_dir, "apex")):
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref not installed (pip install --target baseline/_ref /root/reference)"}))
            return 0
        sys.path.insert(0, ref_dir)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():   # this benchmark is a GPU measurement; say so in one line instead of a CUDA-init traceback
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "no CUDA device on this machine"}))
            return 0
        print("bench.py needs a CUDA device (B200); run it through gpurun or on the GPU box", file=sys.stderr)
        return 2

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        if "MASTER_ADDR" not in os.environ:
            os.environ["MASTER_ADDR"] = "127.0.0.1"
            os.environ.setdefault("MASTER_PORT", str(29500 + (os.getpid() % 2000)))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    from apex_b200.models.llama import make_params, num_params
    from apex_b200.utils.timing import ClockSampler, measured_peaks

    torch.manual_seed(1234)  # identical initial parameters on every rank
    named = make_params(args.config, device=dev, dtype=torch.bfloat16)
    params = [p for _, p in named]
    n_params = num_params(args.config)
    hyper = dict(lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)

    try:
        if args.impl == "ours":
            from apex_b200.contrib.optimizers import DistributedFusedAdam

            opt = DistributedFusedAdam(params, dtype=torch.float32, grad_sync_dtype=torch.bfloat16, param_sync_dtype=torch.bfloat16,
                                       capturable=True, fused_collectives=("auto" if args.fused == "auto" else False), **hyper)
            opt.zero_grad()
            capturable = True
        else:
            from apex.contrib.optimizers.distributed_fused_adam import DistributedFusedAdam as RefDFA

            capturable = True
            try:
                opt = RefDFA(params, dtype=torch.float32, grad_sync_dtype=torch.bfloat16, param_sync_dtype=torch.bfloat16,
                             contiguous_grad_buffer=True, contiguous_param_buffer=True, capturable=True, **hyper)
            except Exception:  # noqa: BLE001
                capturable = False
                opt = RefDFA(params, dtype=torch.float32, grad_sync_dtype=torch.bfloat16, param_sync_dtype=torch.bfloat16,
                             contiguous_grad_buffer=True, contiguous_param_buffer=True, **hyper)
            opt.init_params()
            opt.init_param_buffer()  # documented best practice: re-home parameters before the first step
    except Exception as e:  # noqa: BLE001
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {str(e)[:200]}"}))
            return 0
        raise

    # synthetic gradients (different on every rank), written once straight into the contiguous gradient buffer
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    for p in params:
        opt.grad_buffer_view(p).normal_(0.0, 1e-2, generator=gen)

    def one_step():
        if args.impl == "reference":
            for p in params:  # the reference drops .grad after folding it into the bucket; re-attach the buffer views
                p.grad = opt.grad_buffer_view(p)
        opt.step()

    def lr_tensor():
        lr = opt.param_groups[0]["lr"]
        return lr if torch.is_tensor(lr) else None

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()

    launches0 = getattr(opt, "kernel_launches", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            one_step()
        e1.record()
        torch.cuda.synchronize()
    dist.barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = getattr(opt, "kernel_launches", 0) - launches0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- end to end through the public API: pinned H2D of the step's lr, step, D2H of a result element
    e2e = None
    if not args.no_e2e:
        host_lr = torch.empty(args.steps + 8, dtype=torch.float32).pin_memory()
        for i in range(host_lr.numel()):
            host_lr[i] = hyper["lr"] * (1.0 - 1e-3 * i)
        dev_lr = lr_tensor()
        scratch = torch.zeros(1, dtype=torch.float32, device=dev)
        host_out = torch.empty(1, dtype=torch.float32).pin_memory()
        probe = params[-1].detach().view(-1)

        def e2e_step(i):
            tgt = dev_lr if dev_lr is not None else scratch
            tgt.copy_(host_lr[i:i + 1].view(tgt.shape), non_blocking=True)   # H2D: this step's learning rate
            one_step()
            host_out.copy_(probe[:1].float(), non_blocking=False)             # D2H: one updated parameter (synchronises)
            return float(host_out[0])

        for i in range(3):
            e2e_step(i)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(args.steps):
            e2e_step(3 + i)
        a1.record()
        torch.cuda.synchronize()
        te = torch.tensor([a0.elapsed_time(a1) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": float(te.item()), "unit": "ms/step", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 4,
               "inputs": "per-step learning rate from pinned host memory; gradients are device-resident as produced by backward",
               "lr_on_device": dev_lr is not None}

    if rank == 0:
        pk = measured_peaks()
        D = args.gpus
        hbm_bytes = n_params / D * 28.0
        link_bytes = (D - 1) / D * n_params * 2.0 * 2.0  # per direction: RS pull + AG push
        t_hbm = hbm_bytes / (pk["hbm_gbs"] * 1e9) * 1e3
        t_link = link_bytes / 770e9 * 1e3
        roof = max(t_hbm, t_link)
        cs = clocks.summary()
        out = {
            "metric": "dist_fused_adam_step_ms", "value": ms_max, "unit": "ms/step", "n_gpus": D, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16 grads+params / fp32 optimizer state", "data": "synthetic gradients, random-init weights", "impl": args.impl,
            "config": {"model": args.config, "n_params": n_params, "n_tensors": len(params), "parallelism": f"zero2-dp{D}",
                       "global_batch": None, "seq_len": None, "l2": "working set >> 126 MB L2 (no flush needed)",
                       "fused_collectives": bool(getattr(opt, "fused_collectives", False)),
                       "nvls": bool(getattr(opt, "last_nvls", False)) if args.impl == "ours" else None,
                       "capturable": capturable},
            "clocks": {"sm_mhz": cs["sm_mhz"], "sm_max_mhz": cs["sm_max_mhz"], "reasons": cs["reasons"]},
            "gpu_launches": launches if args.impl == "ours" else None,
            "roofline": {"t_hbm_ms": t_hbm, "t_link_ms": t_link, "bound_ms": roof, "achieved_frac": roof / ms_max,
                         "peaks": pk["source"], "link_gbs": 770.0},
            "params_per_s": n_params / (ms_max * 1e-3),
        }
        if e2e is not None:
            out["e2e"] = e2e
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
