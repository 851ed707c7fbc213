"""``torchsched`` — multi-stream scheduling backend for ``torch.compile``. Reference: apex/contrib/torchsched (2.4k lines): a backend
wrapping Inductor that pins the critical path of the fused-node DAG to the default stream and round-robins the rest over
``TORCH_SCHED_NUM_STREAMS`` side streams with ref-counted CUDA events, plus a pre-grad pass swapping ``F.layer_norm`` for a fused op.

This library does not generate code (explicit kernels, streams and CUDA graphs instead), so ``torch.compile(m, backend="torchsched")``
here applies the same scheduling POLICY to the graph dynamo captures and interprets it: every node is an eager call placed on its
planned stream, with events on the cross-stream edges (:mod:`.scheduler`); ``F.layer_norm`` nodes are rewritten to this library's
fused kernel first. Also exported: :class:`StreamScheduler` (the same fork / join pattern for hand-written step functions) and
:func:`capture_graph` (turn a launch-bound step into a CUDA graph). ``get_backend("inductor")`` still names the stock backend."""
from __future__ import annotations

import os

import torch

from . import config  # noqa: F401
from .scheduler import Plan, ScheduledGraph, plan_graph  # noqa: F401


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


def fused_layer_norm_op(input, normalized_shape, weight=None, bias=None, eps=1e-5):
    """Drop-in for ``F.layer_norm`` (reference torchsched/ops/layer_norm.py): the fused kernel for CUDA fp32 / fp16 / bf16 inputs."""
    if input.is_cuda and input.dtype in (torch.float32, torch.float16, torch.bfloat16):
        # (the package attribute ``normalization.fused_layer_norm`` is the FUNCTION of that name: import from the sub-module)
        from ...normalization.fused_layer_norm import fused_layer_norm, fused_layer_norm_affine

        if weight is not None and bias is not None:
            return fused_layer_norm_affine(input, weight, bias, tuple(normalized_shape), eps)
        if weight is None and bias is None:
            return fused_layer_norm(input, tuple(normalized_shape), eps)
    return torch.nn.functional.layer_norm(input, normalized_shape, weight, bias, eps)


def replace_layer_norm(gm):
    """Pre-grad pass (reference torchsched/passes/pre_grad_passes.py): ``F.layer_norm`` / ``nn.LayerNorm`` nodes -> the fused op."""
    changed = False
    for node in gm.graph.nodes:
        if node.op == "call_function" and node.target is torch.nn.functional.layer_norm:
            node.target = fused_layer_norm_op
            changed = True
        elif node.op == "call_module" and isinstance(gm.get_submodule(node.target), torch.nn.LayerNorm):
            ln = gm.get_submodule(node.target)
            with gm.graph.inserting_before(node):
                w = gm.graph.get_attr(node.target + ".weight") if ln.elementwise_affine else None
                b = gm.graph.get_attr(node.target + ".bias") if ln.elementwise_affine and ln.bias is not None else None
                new = gm.graph.call_function(fused_layer_norm_op, (node.args[0], tuple(ln.normalized_shape), w, b, ln.eps))
            new.meta = dict(node.meta)
            node.replace_all_uses_with(new)
            gm.graph.erase_node(node)
            changed = True
    if changed:
        gm.graph.lint()
        gm.recompile()
    return gm


def _backend(gm, example_inputs, **kwargs):
    gm = replace_layer_norm(gm)
    return ScheduledGraph(gm, num_streams=kwargs.get("num_streams"), cuda_graph=bool(kwargs.get("cuda_graph", False)))


def torchsched(gm, example_inputs, **kwargs):
    """The backend callable itself (reference torchsched/__init__.py:30-43): ``torch.compile(model, backend=torchsched)``."""
    return _backend(gm, example_inputs, **kwargs)


def torchsched_compile(model=None, **kwargs):
    """``torch.compile`` with this backend preselected (reference torchsched/__init__.py:58-81)."""
    kwargs.setdefault("backend", os.environ.get("TORCH_SCHED_DEFAULT_BACKEND", "torchsched"))
    return torch.compile(model, **kwargs) if model is not None else (lambda m: torch.compile(m, **kwargs))


from .backend import DecompositionsWrapper, get_backend  # noqa: E402,F401


def list_backends():
    return ["inductor", "torchsched"]


def set_default_backend(backend: str = "torchsched") -> None:
    """Backend :func:`torchsched_compile` uses when none is given (the reference also swaps it into ``torch.compile`` itself,
    torchsched/__init__.py:44-81; here ``torch.compile`` is left alone)."""
    assert backend in list_backends(), f"Unknown backend {backend}"
    os.environ["TORCH_SCHED_DEFAULT_BACKEND"] = backend


try:  # register so torch.compile(backend="torchsched") resolves
    from torch._dynamo import register_backend

    register_backend(name="torchsched", compiler_fn=_backend)
except Exception:  # noqa: BLE001
    pass

__all__ = ["DecompositionsWrapper", "StreamScheduler", "capture_graph", "set_default_backend", "config", "torchsched", "torchsched_compile", "get_backend", "list_backends",
           "ScheduledGraph", "Plan", "plan_graph", "replace_layer_norm", "fused_layer_norm_op"]
