"""Throughput of GDSFile.save_data / load_data against torch.save / torch.load for a range of sizes
(reference apex/contrib/examples/gpu_direct_storage/benchmark_{save,load}.py).
    python examples/contrib/gpu_direct_storage/benchmark.py [directory] [max_log2_bytes]"""
import os
import sys
import time

import torch

from apex_b200.contrib.gpu_direct_storage import GDSFile


def timed(fn, sync):
    sync()
    t0 = time.perf_counter()
    fn()
    sync()
    return time.perf_counter() - t0


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/tmp"
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    device = "cuda" if torch.cuda.is_available() else "cpu"
    sync = torch.cuda.synchronize if device == "cuda" else (lambda: None)
    print(f"{'bytes':>12} {'gds save':>10} {'torch.save':>10} {'gds load':>10} {'torch.load':>10}   (GB/s)")
    for lg in range(20, top + 1, 2):
        n = (1 << lg) // 4
        x = torch.randn(n, device=device)
        a, b = os.path.join(root, "apex_b200_gds.bin"), os.path.join(root, "apex_b200_torch.pt")

        def gds_save():
            with GDSFile(a, "w") as f:
                f.save_data(x)

        def gds_load():
            with GDSFile(a, "r") as f:
                f.load_data(x)

        rates = [4 * n / timed(fn, sync) / 1e9 for fn in (gds_save, lambda: torch.save(x, b), gds_load, lambda: torch.load(b, map_location=device))]
        print(f"{4 * n:>12} " + " ".join(f"{r:>10.2f}" for r in rates))
        os.remove(a)
        os.remove(b)


if __name__ == "__main__":
    main()
