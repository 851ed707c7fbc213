"""Backend lookup, the convolution-backward decompositions and the compile wrapper. Reference: apex/contrib/torchsched/backend.py:37-348
(``enable_multi_stream_scheduling``, ``convolution_backward_decomp_dwb / _wbd``, ``DecompositionsWrapper``, ``get_backend``).

Two depths of scheduling, chosen by ``config.aot_autograd`` (``TORCH_SCHED_AOT=1``):

* off (default): the forward graph dynamo captured is scheduled over streams; backward is whatever autograd recorded, run by the engine.
* on: AOT autograd traces forward AND backward into ATen graphs — with ``aten.convolution_backward`` split into its data / weight /
  bias gradients in the order the scheme names, so that the three can sit on different streams — and BOTH graphs go through the
  scheduler (this is the granularity the reference schedules at: its ``post_grad_graph_id`` counts these graphs). LayerNorm stays one
  fused node through ``apex_b200::norm_fwd / norm_bwd``. No Inductor and no Triton either way: every node is an eager kernel call."""
from __future__ import annotations

import functools

import torch

from . import config
from .scheduler import ScheduledGraph

aten = torch.ops.aten

__all__ = ["get_backend", "enable_multi_stream_scheduling", "convolution_backward_decomp_dwb", "convolution_backward_decomp_wbd",
           "DecompositionsWrapper"]

# devices on which splitting convolution backward pays (the pieces can overlap on streams); tests add "cpu" to exercise the split
_SPLIT_DEVICES = {"cuda"}


def enable_multi_stream_scheduling(compile_fn):
    """``compile_fn`` with graph lowering patched to the generated multi-stream programs for the duration of the call (reference
    :37-48 patches Inductor's ``GraphLowering`` around ``compile_fx_inner`` the same way)."""
    from .inductor import patch_graph_lowering

    @functools.wraps(compile_fn)
    def wrapper(*args, **kwargs):
        before = config.wrapper_codegen
        patch_graph_lowering(True)
        try:
            return compile_fn(*args, **kwargs)
        finally:
            patch_graph_lowering(before)

    return wrapper


def _conv_backward_piece(which: int, grad_output, input, weight, bias_sizes, stride, padding, dilation, transposed, output_padding, groups):
    mask = [i == which for i in range(3)]
    return aten.convolution_backward(grad_output, input, weight, bias_sizes, stride, padding, dilation, transposed, output_padding, groups,
                                     mask)[which]


def _split_conv_backward(order: str, grad_output, input, weight, bias_sizes, stride, padding, dilation, transposed, output_padding, groups,
                         output_mask):
    """``order`` is a permutation of "dwb": data gradient, weight gradient, bias gradient — the order in which the three independent
    pieces are issued (= the order the scheduler sees them, which decides who gets the caller's stream first)."""
    if not output_mask[2] or grad_output.device.type not in _SPLIT_DEVICES:
        return NotImplemented        # nothing to overlap with: keep the single fused call
    geometry = (bias_sizes, stride, padding, dilation, transposed, output_padding, groups)
    out = {}
    for piece in order:
        if piece == "d":
            out["d"] = _conv_backward_piece(0, grad_output, input, weight, *geometry) if output_mask[0] else None
        elif piece == "w":
            out["w"] = _conv_backward_piece(1, grad_output, input, weight, *geometry) if output_mask[1] else None
        else:
            out["b"] = aten.sum(grad_output, [0] + list(range(2, grad_output.dim())))
    return out["d"], out["w"], out["b"]


def convolution_backward_decomp_dwb(grad_output, input, weight, bias_sizes, stride, padding, dilation, transposed, output_padding, groups,
                                    output_mask):
    """dgrad, then wgrad, then bgrad (reference :51-113): the next layer's backward waits for dgrad, so it goes first."""
    return _split_conv_backward("dwb", grad_output, input, weight, bias_sizes, stride, padding, dilation, transposed, output_padding, groups,
                                output_mask)


def convolution_backward_decomp_wbd(grad_output, input, weight, bias_sizes, stride, padding, dilation, transposed, output_padding, groups,
                                    output_mask):
    """wgrad, bgrad, then dgrad (reference :116-178)."""
    return _split_conv_backward("wbd", grad_output, input, weight, bias_sizes, stride, padding, dilation, transposed, output_padding, groups,
                                output_mask)


class DecompositionsWrapper:
    """The object ``get_backend("torchsched")`` returns and ``torch.compile(backend=...)`` calls with (graph module, example inputs).
    Reference :181-262 subclasses torch's Inductor wrapper and adds the decomposition table; this one owns the whole pipeline:
    pre-grad passes -> (optionally) AOT autograd with the decomposition table -> :class:`ScheduledGraph` per graph."""

    def __init__(self, mode: str = "default", options: dict | None = None, dynamic: bool = False, decompositions: dict | None = None) -> None:
        self.mode, self.dynamic = mode, dynamic
        self.config = dict(options or {})
        self.decompositions = dict(decompositions or {})
        self.graphs: list = []       # every ScheduledGraph built through this wrapper, in compilation order

    def __eq__(self, rhs: object) -> bool:
        return (isinstance(rhs, DecompositionsWrapper) and (self.mode, self.dynamic, self.config, self.decompositions)
                == (rhs.mode, rhs.dynamic, rhs.config, rhs.decompositions))

    __hash__ = object.__hash__

    def _schedule(self, gm, example_inputs=None, wrapper_codegen=None):
        sg = ScheduledGraph(gm, num_streams=self.config.get("num_streams"), cuda_graph=bool(self.config.get("cuda_graph", False)))
        if wrapper_codegen is not None:
            sg.wrapper_codegen = wrapper_codegen
        self.graphs.append(sg)
        return sg

    def __call__(self, model_, inputs_, *args, **kwargs):
        from .passes import pre_grad_custom_pass

        use_aot = self.config.get("aot_autograd", config.aot_autograd)
        if config.enable_pre_grad_pass:
            pre_grad_custom_pass(model_.graph, traceable=bool(use_aot))
        if not use_aot:
            return self._schedule(model_, inputs_)
        from functorch.compile import make_boxed_func
        from torch._dynamo.backends.common import aot_autograd

        codegen = bool(config.wrapper_codegen)     # the backward graph is compiled lazily, at the first backward: it keeps today's choice

        def compiler(gm, example_inputs):
            return make_boxed_func(self._schedule(gm, example_inputs, codegen))

        return aot_autograd(fw_compiler=compiler, bw_compiler=compiler, decompositions=self.decompositions)(model_, inputs_)


def get_backend(backend: str = "torch", scheme: str = "dwb"):
    """``"torch"`` -> the stock Inductor backend (by name); ``"torchsched"`` -> a :class:`DecompositionsWrapper` whose convolution
    backward is split in ``scheme`` order ("dwb" or "wbd", reference :265-348)."""
    if backend not in ("torch", "inductor", "torchsched"):
        raise ValueError(f"Unknown compilation {backend=}")
    if scheme not in ("dwb", "wbd"):
        raise ValueError(f"Invalid {scheme=}, use scheme=dwb or wbd instead")
    if backend != "torchsched":
        return "inductor"
    decomp = convolution_backward_decomp_dwb if scheme == "dwb" else convolution_backward_decomp_wbd
    return DecompositionsWrapper(mode="default", options={}, dynamic=False, decompositions={aten.convolution_backward.default: decomp})
