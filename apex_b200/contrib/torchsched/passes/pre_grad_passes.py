"""Pre-grad graph passes (reference apex/contrib/torchsched/passes/pre_grad_passes.py:24-100): a registry of rewrites applied to the captured
graph before it is scheduled; ``replace_layer_norm`` is the one the reference ships."""
from __future__ import annotations

import torch

from .. import replace_layer_norm

_PASSES: dict = {"layer_norm": replace_layer_norm}


def register_pattern(name: str, rewrite) -> None:
    """``rewrite(gm) -> gm``; applied by :func:`run_pre_grad_pass` in registration order."""
    _PASSES[name] = rewrite


def run_pre_grad_pass(gm: torch.fx.GraphModule) -> torch.fx.GraphModule:
    for rewrite in _PASSES.values():
        gm = rewrite(gm)
    return gm


def pre_grad_custom_pass(graph: torch.fx.Graph) -> None:
    """Graph-level entry point (the hook signature Inductor's ``pre_grad_custom_pass`` config expects): rewrites ``graph`` in place."""
    from .. import fused_layer_norm_op

    for node in graph.nodes:
        if node.op == "call_function" and node.target is torch.nn.functional.layer_norm:
            node.target = fused_layer_norm_op
    graph.lint()


__all__ = ["register_pattern", "replace_layer_norm", "run_pre_grad_pass", "pre_grad_custom_pass"]
