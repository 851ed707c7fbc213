"""Exposed communication of a ZeRO-2 training step on Llama-shaped FFN blocks built from this library's FusedDense modules.

    python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node N benchmarks/bench_llama_block.py

Per configuration: ms / iteration of  zero_grad -> forward -> backward -> optimizer.step()  (CUDA events, max over ranks), for
  compute   : forward + backward only (no optimizer): the floor
  serial    : DistributedFusedAdam(overlap_grad_sync=False): every collective exposed inside step()
  overlap   : overlap_grad_sync=True (per-bucket reduce-scatter during backward) + overlap_param_sync=True (Adam + parameter push on
              the side stream; the next forward's GEMMs acquire per-bucket flags tile by tile — the all-gather fused into the GEMM)
exposed = (iteration - compute - local Adam time at HBM speed) ; reported as a fraction of the iteration."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--ffn", type=int, default=14336)
    ap.add_argument("--tokens", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=8)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from apex_b200.contrib.optimizers import DistributedFusedAdam
    from apex_b200.fused_dense import FusedDense
    from apex_b200.normalization import FusedRMSNorm
    from apex_b200.parallel.param_sync import attach_param_sync_hooks

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.norm = FusedRMSNorm(args.hidden)
            self.up, self.gate, self.down = FusedDense(args.hidden, args.ffn, bias=False), FusedDense(args.hidden, args.ffn, bias=False), \
                FusedDense(args.ffn, args.hidden, bias=False)

        def forward(self, x):
            h = self.norm(x)
            return x + self.down(torch.nn.functional.silu(self.gate(h)) * self.up(h))

    def build():
        torch.manual_seed(0)
        return torch.nn.Sequential(*[Block() for _ in range(args.layers)]).to(dev, torch.bfloat16)

    x = torch.randn(args.tokens, args.hidden, device=dev, dtype=torch.bfloat16)

    def run(model, opt):
        def it():
            if opt is not None:
                opt.zero_grad(set_to_none=True)
            else:
                model.zero_grad(set_to_none=True)
            model(x).float().pow(2).mean().backward()
            if opt is not None:
                opt.step()
        for _ in range(3):
            it()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            it()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / args.iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    out = {"world": world, "layers": args.layers, "hidden": args.hidden, "ffn": args.ffn, "tokens_per_gpu": args.tokens}
    m = build()
    out["n_params"] = sum(p.numel() for p in m.parameters())
    out["compute_ms"] = run(m, None)
    del m
    m = build()
    out["serial_ms"] = run(m, DistributedFusedAdam(m.parameters(), lr=1e-4, overlap_grad_sync=False, process_group=dist.new_group(list(range(world)))))
    del m
    m = build()
    opt = DistributedFusedAdam(m.parameters(), lr=1e-4, overlap_grad_sync=True, overlap_param_sync=True, process_group=dist.new_group(list(range(world))))
    attach_param_sync_hooks(m)
    out["overlap_ms"] = run(m, opt)
    opt.param_sync()
    adam_ms = out["n_params"] / world * 28.0 / 6.5e12 * 1e3          # local Adam at ~HBM speed: not communication
    for k in ("serial", "overlap"):
        out[f"{k}_exposed_comm_ms"] = max(0.0, out[f"{k}_ms"] - out["compute_ms"] - adam_ms)
        out[f"{k}_exposed_frac"] = out[f"{k}_exposed_comm_ms"] / out[f"{k}_ms"]
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
