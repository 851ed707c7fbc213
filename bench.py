#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): DistributedFusedAdam step time over the Llama-3-8B parameter set.

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--config llama3-8b]

One process per GPU (the driver launches N>1 through torch.distributed.run). fp32 optimizer state, bf16 gradients, bf16 parameters
(the configuration BASELINE.md section 2 row 4 names). Strong scaling: the parameter set is fixed, each rank owns 1/N of the state.

Both arms run EXACTLY the same per-step sequence through their public API:

    opt.zero_grad()                      # re-attaches p.grad = view of the contiguous gradient buffer, zeroes the buffer
    p.grad.copy_(synthetic) for every p  # stands in for backward: different values on every rank, written in place
    opt.step()                           # reduce-scatter + Adam on the fp32 shard + all-gather of the new bf16 parameters

* ``value`` / ``ms_per_step``: the metric BASELINE names — the step() call alone, one CUDA-event pair per step on the launching
  stream, summed over the K timed steps, MAX over ranks.  ``sequence_ms_per_step`` is the whole sequence (zero_grad + refill + step)
  between one event pair, same K steps.
* ``e2e``: a real training step through the public API: tokens / labels copied host->device from pinned memory every step, forward
  + backward of the SAME plain-PyTorch Llama (bench_common.llama_loss, torch.nn.functional only) over these parameters — backward
  writes the gradients in place into the optimizer's gradient buffer and triggers whatever overlap hooks the optimizer installs —
  then opt.step(), then a device->host read of the loss.
The working set (>= 44 GB per GPU at N=8, 128 GB at N=1) is far larger than the 126 MB L2, so no L2 flush is needed between steps.
Clocks are sampled with nvidia-smi during the timed region.

--impl reference runs the UNMODIFIED reference (baseline/_ref, NCCL + its own kernels); that process imports only torch,
bench_common and the reference's ``apex`` — nothing from this repo's package.
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
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="llama3-8b")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-seq", type=int, default=1024, help="tokens per sequence of the end-to-end training step")
    ap.add_argument("--e2e-batch", type=int, default=1, help="sequences per GPU of the end-to-end training step")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default: --steps)")
    ap.add_argument("--fused", default="auto", choices=["auto", "off"], help="ours: in-kernel collectives (auto) or the NCCL path (off)")
    ap.add_argument("--overlap", default="off", choices=["on", "off"],
                    help="ours: overlap_grad_sync hooks during backward (e2e). Off by default: measured SLOWER end to end on this workload "
                         "(4 GPUs: 103.4 ms with the per-bucket reduce-scatter launches under backward, 96.6 ms with the one-kernel step)")
    ap.add_argument("--step-in-backward", default="off", choices=["on", "off"],
                    help="ours: overlap_step_with_backward=True (each bucket's whole step runs from the gradient hook during backward)")
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

    from bench_common import ClockSampler, llama_loss, make_params, measured_peaks, num_params, CONFIGS

    K, W = args.steps, max(args.warmup, 3)
    torch.manual_seed(1234)  # identical initial parameters on every rank
    named = make_params(args.config, device=dev, dtype=torch.bfloat16)
    params = [p for _, p in named]
    P = dict(named)
    n_params = num_params(args.config)
    hyper = dict(lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)

    try:
        if args.impl == "ours":
            from apex_b200.contrib.optimizers import DistributedFusedAdam

            opt = DistributedFusedAdam(params, dtype=torch.float32, grad_sync_dtype=torch.bfloat16, param_sync_dtype=torch.bfloat16,
                                       capturable=True, fused_collectives=("auto" if args.fused == "auto" else False),
                                       overlap_grad_sync=(args.overlap == "on"),
                                       overlap_step_with_backward=(args.step_in_backward == "on"), **hyper)
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

    # synthetic gradient source (different on every rank): one pool as large as the largest parameter, copied into p.grad every step
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    pool = torch.empty(max(p.numel() for p in params), device=dev, dtype=torch.bfloat16).normal_(0.0, 1e-2, generator=gen)

    def refill():
        for p in params:
            g = p.grad if p.grad is not None else opt.grad_buffer_view(p)
            g.copy_(pool[:p.numel()].view_as(p))
            if p.grad is None:
                p.grad = g

    def sequence(ev=None):
        opt.zero_grad()
        refill()
        if ev is not None:
            ev[0].record()
        opt.step()
        if ev is not None:
            ev[1].record()

    for _ in range(W):
        sequence()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()

    # ---- (1) the metric: step() alone, one event pair per step; (2) the whole sequence between one pair -- same K steps
    launches0 = getattr(opt, "kernel_launches", 0)
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        s0.record()
        for i in range(K):
            sequence(pairs[i])
        s1.record()
        torch.cuda.synchronize()
    dist.barrier()
    launches = getattr(opt, "kernel_launches", 0) - launches0
    step_ms = sum(a.elapsed_time(b) for a, b in pairs) / K
    seq_ms = s0.elapsed_time(s1) / K
    t = torch.tensor([step_ms, seq_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, seq_max = float(t[0].item()), float(t[1].item())

    # ---- end to end through the public API: pinned H2D of tokens / labels, forward + backward (gradients land in the optimizer's
    # buffer in place, overlap hooks fire), step, D2H of the loss
    e2e = None
    if not args.no_e2e:
        Ke = args.e2e_steps or K
        B, S = args.e2e_batch, args.e2e_seq
        vocab = CONFIGS[args.config]["vocab"]
        g2 = torch.Generator().manual_seed(99 + rank)
        host_tok = torch.randint(0, vocab, (Ke + 3, B, S), generator=g2).pin_memory()
        host_lab = torch.randint(0, vocab, (Ke + 3, B, S), generator=g2).pin_memory()
        host_loss = torch.empty(1, dtype=torch.float32).pin_memory()
        e2e_launch0 = getattr(opt, "kernel_launches", 0)
        try:
            def e2e_step(i):
                tok = host_tok[i].to(dev, non_blocking=True)
                lab = host_lab[i].to(dev, non_blocking=True)
                opt.zero_grad(set_to_none=True)   # same call in both arms
                loss = llama_loss(P, tok, lab, args.config)
                loss.backward()
                opt.step()
                host_loss.copy_(loss.detach().reshape(1), non_blocking=False)   # D2H: the step's loss (synchronises)
                return float(host_loss[0])

            losses = [e2e_step(i) for i in range(3)]
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(Ke):
                losses.append(e2e_step(3 + i))
            a1.record()
            torch.cuda.synchronize()
            te = torch.tensor([a0.elapsed_time(a1) / Ke], device=dev, dtype=torch.float64)
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e = {"value": float(te.item()), "unit": "ms/step", "h2d_bytes_per_step": 2 * B * S * 8, "d2h_bytes_per_step": 4,
                   "steps": Ke, "what": "H2D tokens+labels (pinned) -> zero_grad(set_to_none=True) -> plain-PyTorch Llama forward+backward (gradients written in "
                   "place into the optimizer's buffer) -> optimizer.step() -> D2H loss", "batch_per_gpu": B, "seq_len": S,
                   "tokens_per_s": args.gpus * B * S / (float(te.item()) * 1e-3), "loss_first": losses[0], "loss_last": losses[-1],
                   "gpu_launches": (getattr(opt, "kernel_launches", 0) - e2e_launch0) if args.impl == "ours" else None,
                   "step_in_backward": bool(getattr(opt, "overlap_step_with_backward", False)),
                   "overlap_grad_sync": bool(getattr(opt, "overlap_grad_sync", False)) if args.impl == "ours" else "reference default (True)"}
        except torch.OutOfMemoryError as e:   # noqa: PERF203
            e2e = {"unavailable": f"out of memory in the end-to-end step: {str(e)[:120]}"}

    if rank == 0:
        pk = measured_peaks()
        D = args.gpus
        nvls = bool(getattr(opt, "last_nvls", False)) if args.impl == "ours" else None
        hbm_bytes = n_params / D * 28.0
        S_bytes = n_params * 2.0
        # bytes per direction per GPU: P2P pulls (D-1)/D of the gradients and pushes (D-1)/D of the parameters in BOTH directions;
        # through the switch (NVLS) every rank sends its whole gradient buffer once and receives 1/D reduced, and the reverse for parameters
        link_bytes = 0.0 if D == 1 else ((S_bytes + S_bytes / D) if nvls else 2.0 * (D - 1) / D * S_bytes)
        t_hbm = hbm_bytes / (pk["hbm_gbs"] * 1e9) * 1e3
        t_link = link_bytes / (pk["link_gbs"] * 1e9) * 1e3
        roof = max(t_hbm, t_link)
        cs = clocks.summary()
        out = {
            "metric": "dist_fused_adam_step_ms", "value": ms_max, "unit": "ms/step", "n_gpus": D, "steps": K,
            "warmup": W, "ms_per_step": ms_max, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16 grads+params / fp32 optimizer state", "data": "synthetic gradients / tokens, random-init weights", "impl": args.impl,
            "timing": "sum of per-step CUDA-event pairs around optimizer.step(); zero_grad() and the gradient refill run between the pairs "
                      "(identical in both arms)", "sequence_ms_per_step": seq_max,
            "config": {"model": args.config, "n_params": n_params, "n_tensors": len(params), "parallelism": f"zero2-dp{D}",
                       "global_batch": D * args.e2e_batch, "seq_len": args.e2e_seq, "l2": "working set >> 126 MB L2 (no flush needed)",
                       "fused_collectives": bool(getattr(opt, "fused_collectives", False)), "nvls": nvls,
                       "capturable": capturable},
            "clocks": {"sm_mhz": cs["sm_mhz"], "sm_max_mhz": cs["sm_max_mhz"], "reasons": cs["reasons"]},
            "gpu_launches": launches if args.impl == "ours" else None,
            "roofline": {"t_hbm_ms": t_hbm, "t_link_ms": t_link, "bound_ms": roof, "achieved_frac": min(1.0, roof / ms_max),
                         "peaks": pk["source"], "link_gbs": pk["link_gbs"], "link_source": pk["link_source"],
                         "link_bytes_per_direction": link_bytes},
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
