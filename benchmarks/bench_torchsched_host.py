"""Host-side cost per call of the two torchsched execution modes (interpreter vs generated program) and of FX's own generated forward, on
tiny CPU tensors so that the kernels cost next to nothing and what is left is Python overhead per node. Needs no GPU.
usage: python benchmarks/bench_torchsched_host.py [--iters 2000]"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from apex_b200.contrib import torchsched as ts  # noqa: E402
from apex_b200.contrib.torchsched.inductor import lower_graph  # noqa: E402


class Net(nn.Module):
    def __init__(self, depth=8):
        super().__init__()
        self.main = nn.ModuleList(nn.Linear(16, 16) for _ in range(depth))
        self.side = nn.ModuleList(nn.Linear(16, 16) for _ in range(depth))
        self.norm = nn.LayerNorm(16)

    def forward(self, x):
        h = x
        outs = []
        for a, b in zip(self.main, self.side):
            h = F.gelu(a(h))
            outs.append(torch.tanh(b(x)))
        return self.norm(h + sum(outs))


def timeit(fn, args, iters):
    for _ in range(50):
        fn(*args)
    t0 = time.perf_counter()
    for _ in range(iters):
        fn(*args)
    return (time.perf_counter() - t0) / iters * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2000)
    a = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(1)
    m, x = Net(), torch.randn(2, 16)
    cap = {}

    def backend(gm, ex):
        sg = ts._backend(gm, ex)
        cap["sg"] = sg

        def run(*args):
            cap["args"] = args
            return sg(*args)
        return run

    with torch.no_grad():
        torch.compile(m, backend=backend)(x)
        sg, args = cap["sg"], cap["args"]
        n_nodes = len(sg.plan.order)
        program = lower_graph(sg.gm, multi_stream=False)
        res = {"nodes": n_nodes,
               "interpreter_us": timeit(sg, args, a.iters),
               "generated_program_us": timeit(program, args, a.iters),
               "fx_forward_us": timeit(sg.gm.forward, args, a.iters),
               "eager_module_us": timeit(m, (x,), a.iters)}
    res["interpreter_over_program"] = res["interpreter_us"] / res["generated_program_us"]
    res["per_node_saving_us"] = (res["interpreter_us"] - res["generated_program_us"]) / n_nodes
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}))


if __name__ == "__main__":
    main()
