"""Switch between the two ways a scheduled graph is executed, and the code dump. Reference: apex/contrib/torchsched/inductor/graph.py:31-125
— there ``patch_graph_lowering`` swaps Inductor's ``GraphLowering.codegen`` for one that schedules with ``MultiCudaStreamScheduler`` and
writes through ``MultiStreamWrapperCodegen``; ``TORCH_SCHED_DUMP_CODE`` additionally writes the generated wrapper (and, with
``+inductor,``, the stock one) per graph.

Here the thing being lowered is the FX graph itself: patched, :class:`..scheduler.ScheduledGraph` runs the generated multi-stream
program of :mod:`.scheduler`; unpatched, it interprets the graph node by node. Same plan, same kernels, same results."""
from __future__ import annotations

from pathlib import Path

import torch.fx as fx

from .. import config as torchsched_config
from .scheduler import MultiCudaStreamScheduler

__all__ = ["patch_graph_lowering", "lower_graph"]


def _torchsched_codegen(gm: fx.GraphModule, num_streams=None, graph_id: int = 0, multi_stream=None):
    """(callable, program text) of the multi-stream program of ``gm``."""
    sched = MultiCudaStreamScheduler(gm, num_streams=num_streams, graph_id=graph_id, multi_stream=multi_stream)
    src = sched.codegen()
    return sched.compile(), src, sched


def _single_stream_codegen(gm: fx.GraphModule, graph_id: int = 0):
    """The baseline a dump compares against: what FX itself generates for the graph (one stream, program order)."""
    return gm.forward, gm.code


def _dump(graph_id: int, backend: str, text: str) -> None:
    out = Path(torchsched_config.dump_code_dir) / backend
    out.mkdir(parents=True, exist_ok=True)
    (out / f"graph_{graph_id}_wrapper_code.py").write_text(text)


def lower_graph(gm: fx.GraphModule, num_streams=None, graph_id: int = 0, multi_stream=None):
    """Generated program of ``gm`` (what a patched :class:`ScheduledGraph` calls); writes the dump when one is configured."""
    fn, src, sched = _torchsched_codegen(gm, num_streams, graph_id, multi_stream)
    if torchsched_config.dump_code_dir:
        for backend in torchsched_config.dump_code_backends:
            if backend == "torchsched":
                _dump(graph_id, backend, src)
            elif backend == "inductor":
                _dump(graph_id, backend, _single_stream_codegen(gm, graph_id)[1])
            else:
                raise ValueError(f"Unknown {backend=} from {torchsched_config.dump_code_backends=}")
    fn.scheduler = sched
    return fn


def patch_graph_lowering(patch: bool = True) -> None:
    """``True``: graphs compiled from now on run their generated multi-stream program; ``False``: back to the interpreter."""
    torchsched_config.wrapper_codegen = bool(patch)
