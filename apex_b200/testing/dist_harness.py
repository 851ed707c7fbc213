"""Multi-process test harness (the reference's NcclDistributedTestBase analogue, apex/distributed_testing/distributed_test_base.py:25-131):
spawns ``world_size`` processes on ONE node, rendezvous over a file store on 127.0.0.1-free file:// init, runs ``fn(rank, world, *args)``.
Backend: nccl when every rank can own a GPU, else gloo (CPU) — so host-side logic is testable without a GPU."""
from __future__ import annotations

import os
import tempfile
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _entry(rank, world, backend, init_file, fn, args, errq):
    try:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(rank % torch.cuda.device_count())
        dist.init_process_group(backend=backend, init_method=f"file://{init_file}", rank=rank, world_size=world)
        try:
            fn(rank, world, *args)
            dist.barrier()
        finally:
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        errq.put((rank, traceback.format_exc()))
        raise


def run_distributed(fn, world_size: int, *args, backend: str | None = None, timeout: float = 240.0):
    """Run ``fn(rank, world_size, *args)`` in ``world_size`` processes; re-raises the first failure with its traceback."""
    if backend is None:
        backend = "nccl" if (torch.cuda.is_available() and torch.cuda.device_count() >= world_size) else "gloo"
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rdzv")
        procs = [ctx.Process(target=_entry, args=(r, world_size, backend, init_file, fn, args, errq), daemon=True) for r in range(world_size)]
        for p in procs:
            p.start()
        # Poll instead of joining one by one: when a rank dies its peers usually sit in a collective until their own time-out, so the
        # first failure ends the run after a short grace period (GPU box minutes are budgeted) and the overall deadline is a single one.
        import time

        deadline = time.monotonic() + timeout
        failed_at = None
        while any(p.is_alive() for p in procs):
            now = time.monotonic()
            if failed_at is None and any((not p.is_alive()) and p.exitcode != 0 for p in procs):
                failed_at = now
            if now > deadline or (failed_at is not None and now - failed_at > 10.0):
                break
            time.sleep(0.05)
        bad = [p for p in procs if p.is_alive() or p.exitcode != 0]
        for p in procs:
            if p.is_alive():
                p.terminate()
        for p in procs:
            p.join(5.0)
        if bad:
            msg = ""
            while not errq.empty():
                r, tb = errq.get()
                msg += f"\n--- rank {r} ---\n{tb}"
            raise RuntimeError(f"distributed test failed ({len(bad)} rank(s)){msg or ' (timeout / crash without traceback)'}")
