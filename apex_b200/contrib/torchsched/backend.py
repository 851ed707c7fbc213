"""Import-path parity with apex/contrib/torchsched/backend.py: backend lookup and the decorator that turns a compile function into a
multi-stream one. The Inductor-specific pieces of the reference (convolution-backward decompositions, wrapper code generation) have
no counterpart: graphs are interpreted on streams (see :mod:`.scheduler`)."""
from __future__ import annotations

import functools

from . import get_backend, torchsched  # noqa: F401


def enable_multi_stream_scheduling(compile_fn):
    """``compile_fn(gm, example_inputs, ...)`` -> the same call routed through the multi-stream scheduler (reference :37-48)."""

    @functools.wraps(compile_fn)
    def wrapper(gm, example_inputs, *args, **kwargs):
        return torchsched(gm, example_inputs, **{k: v for k, v in kwargs.items() if k in ("num_streams", "cuda_graph")})

    return wrapper
