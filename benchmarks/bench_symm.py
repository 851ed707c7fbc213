"""Symmetric-heap micro-benchmarks.

    python -m torch.distributed.run --standalone --local-addr 127.0.0.1 --nproc-per-node N benchmarks/bench_symm.py [--mb 256]

Per op and unroll depth: GB/s moved by ONE rank's kernel while every rank runs the same probe (device-timed with CUDA events, max over ranks),
plus the latency of the in-kernel barrier. Ops: peer read / peer write over NVLink P2P (neighbour rank), multimem.ld_reduce / multimem.st through
the NVSwitch (NVLS). The numbers bound what the fused ZeRO step (csrc/dist_adam.cu) can reach: SURVEY.md section 7.2 step 5."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apex_b200 import _lib  # noqa: E402
from apex_b200.parallel.symmetric import SignalPad, SymmetricMemory  # noqa: E402

_lib.declare("ab_symm_bench", "i p p l i i p p")


def timed(fn, iters, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ctas", type=int, nargs="*", default=None)
    ap.add_argument("--unroll", type=int, nargs="*", default=None)
    ap.add_argument("--write-peaks", action="store_true", help="store the best observed link rate in profiles/results/link_peaks.json")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    nbytes = args.mb << 20
    mem = SymmetricMemory(nbytes, multicast=True, tag="bw")
    mem2 = SymmetricMemory(nbytes, multicast=True, tag="bw2")
    local_buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    mem.buffer.view(torch.float32).fill_(1.0)
    sink = torch.zeros(1, device=dev)
    pad = SignalPad.get(None, dev)
    peer = mem.peer_ptrs[(rank + 1) % world]
    stream = lambda: _lib.stream_ptr(dev)  # noqa: E731
    rows = []
    # (name, op, src, dst, bytes moved by this rank's kernel, link bytes per direction on THIS GPU per payload byte)
    probes = [("peer_read", 0, peer, local_buf.data_ptr(), nbytes, 1.0), ("peer_write", 1, local_buf.data_ptr(), peer, nbytes, 1.0)]
    if mem.has_multicast and mem2.has_multicast:
        per_rank = nbytes // world // 16 * 16       # every rank reduces / broadcasts its own slice, as the ZeRO kernel does
        off = rank * per_rank
        # every rank's ld_reduce pulls its slice out of ALL GPUs: this GPU's egress = world x slice; multimem.st: ingress = world x slice
        probes += [("nvls_ld_reduce", 2, mem.mc_ptr + off, local_buf.data_ptr(), per_rank, float(world)),
                   ("nvls_st", 3, local_buf.data_ptr(), mem.mc_ptr + off, per_rank, float(world)),
                   ("nvls_ld_reduce+st", 4, mem.mc_ptr + off, mem2.mc_ptr + off, per_rank, float(world) + 1.0)]
    best = {}
    for name, op, src, dst, size, link_factor in probes:
        for ctas in (args.ctas or [74, 148, 296, 592]):
            for unroll in (args.unroll or [2, 8]):
                ms = timed(lambda: _lib.fn("ab_symm_bench")(op, src, dst, size, ctas, unroll, sink.data_ptr(), stream()), args.iters, dev)
                link = size * link_factor / ms / 1e6
                rows.append({"op": name, "ctas": ctas, "unroll": unroll, "ms": ms, "payload_GBps": size / ms / 1e6,
                             "link_GBps_per_direction": link, "world": world})
                best[name] = max(best.get(name, 0.0), link)
    ms = timed(lambda: pad.barrier(channel=50), 200, dev)
    rows.append({"op": "barrier", "us": ms * 1e3, "world": world})
    if rank == 0:
        for r in rows:
            print(json.dumps(r))
        print(json.dumps({"summary": best, "world": world, "mb": args.mb}))
        if args.write_peaks:
            out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "results", "link_peaks.json")
            json.dump({"link_gbs": max(best.values()), "per_op_best_GBps": best, "world": world, "mb": args.mb,
                       "how": "benchmarks/bench_symm.py: best sustained bytes per direction on one GPU's links over the probes "
                              "(P2P read / write, multimem.ld_reduce, multimem.st, ld_reduce+st)"}, open(out, "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
