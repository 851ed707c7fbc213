"""Tracing helpers. The reference has no tracing framework — only ad-hoc NVTX ranges and cudaProfilerStart/Stop brackets in its
examples and tests (examples/imagenet/main_amp.py:317-392, tests/L0/run_mlp/test_mlp.py:172,196); these are the same tools as a
small API, plus the CUDA-event timers of :mod:`apex_b200.utils.timing`.

    with nvtx_range("optimizer.step()"): opt.step()
    prof = ProfilerWindow(start=20, steps=10)          # cudaProfilerStart at iteration 20, Stop after 10 (for nsys / ncu --profile-from-start off)
    for i, batch in enumerate(loader): prof.step(i); ...
    annotate_modules(model)                            # NVTX range per leaf module forward (and backward via autograd hooks)
"""
from __future__ import annotations

import contextlib
import functools

import torch


def _nvtx():
    return torch.cuda.nvtx if torch.cuda.is_available() else None


@contextlib.contextmanager
def nvtx_range(name: str):
    n = _nvtx()
    if n is not None:
        n.range_push(name)
    try:
        yield
    finally:
        if n is not None:
            n.range_pop()


def annotate(name: str | None = None):
    """Decorator: run the function inside an NVTX range (default: its qualified name)."""

    def deco(fn):
        label = name or fn.__qualname__

        @functools.wraps(fn)
        def wrapper(*a, **k):
            with nvtx_range(label):
                return fn(*a, **k)

        return wrapper

    return deco


class ProfilerWindow:
    """cudaProfilerStart / cudaProfilerStop around iterations [start, start + steps) — the ``--prof N`` switch of the reference example."""

    def __init__(self, start: int = -1, steps: int = 10):
        self.start, self.stop, self.on = start, start + steps, False

    def step(self, iteration: int) -> None:
        if self.start < 0 or not torch.cuda.is_available():
            return
        if iteration == self.start and not self.on:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            self.on = True
        elif iteration == self.stop and self.on:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            self.on = False


def annotate_modules(model: torch.nn.Module, leaf_only: bool = True):
    """Wrap every (leaf) module's forward in an NVTX range named after the module path; returns the handles (call .remove())."""
    handles = []
    for name, m in model.named_modules():
        if leaf_only and any(True for _ in m.children()):
            continue
        label = f"{name or 'model'}:{type(m).__name__}"
        handles.append(m.register_forward_pre_hook(lambda mod, inp, _l=label: _nvtx() and _nvtx().range_push(_l)))
        handles.append(m.register_forward_hook(lambda mod, inp, out: _nvtx() and _nvtx().range_pop()))
    return handles
