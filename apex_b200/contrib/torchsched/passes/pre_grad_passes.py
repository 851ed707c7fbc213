"""Pre-grad graph passes: named (pattern, replacement) rewrites applied to the captured graph before it is scheduled. Reference:
apex/contrib/torchsched/passes/pre_grad_passes.py:24-107 (``register_pattern``, ``run_pre_grad_pass`` over ``torch.fx.replace_pattern``,
``pre_grad_custom_pass`` driven by ``config.pre_grad_pass_options``; the one pass it ships swaps ``F.layer_norm`` for a cuDNN-graph op).

A pattern here is a single callable (``torch.nn.functional.layer_norm``): every ``call_function`` node with that target is re-pointed at
the replacement, with its arguments normalised to the pattern's signature — positional or keyword, defaults filled in — so the
replacement always receives them positionally in declaration order (the reference has to patch kwargs into args by hand for the same
reason, :79-86)."""
from __future__ import annotations

import inspect
import logging

import torch

from .. import config, fused_layer_norm_op

__all__ = ["pre_grad_custom_pass", "register_pattern", "replace_layer_norm", "replace_layer_norm_traceable", "run_pre_grad_pass",
           "PRE_GRAD_PASS_PATTERNS"]

# pass name -> (pattern, replacement, replacement to use when the graph is going to be traced again by AOT autograd)
PRE_GRAD_PASS_PATTERNS: dict = {}
counters: dict = {}


def register_pattern(name: str, pattern, replacement, traceable_replacement=None) -> None:
    if name in PRE_GRAD_PASS_PATTERNS:
        raise ValueError(f"pre-grad pass {name!r} is already registered")
    PRE_GRAD_PASS_PATTERNS[name] = (pattern, replacement, traceable_replacement or replacement)


def replace_layer_norm(x, normalized_shape, weight, bias, eps):
    """``F.layer_norm`` -> this library's fused LayerNorm kernel (eager autograd path)."""
    return fused_layer_norm_op(x, normalized_shape, weight, bias, eps)


def replace_layer_norm_traceable(x, normalized_shape, weight, bias, eps):
    """Same, through the ``apex_b200::norm_fwd`` custom op: AOT autograd can trace through it (fake kernel + registered backward) and
    the fused kernels stay single nodes of the forward and backward graphs."""
    from ..ops.layer_norm import layer_norm

    return layer_norm(x, list(normalized_shape), weight, bias, eps)[0]


register_pattern("fused_layer_norm", torch.nn.functional.layer_norm, replace_layer_norm, replace_layer_norm_traceable)


def run_pre_grad_pass(name: str, graph: torch.fx.Graph, pattern, replacement) -> int:
    """Re-point every call of ``pattern`` in ``graph`` at ``replacement``; returns how many were rewritten."""
    try:
        sig = inspect.signature(pattern)
    except (TypeError, ValueError):
        sig = None
    n = 0
    for node in list(graph.nodes):
        if node.op != "call_function" or node.target is not pattern:
            continue
        if sig is not None:
            try:
                bound = sig.bind(*node.args, **node.kwargs)
            except TypeError:
                continue                 # a call the pattern's own signature rejects: leave it alone
            bound.apply_defaults()
            node.args, node.kwargs = tuple(bound.arguments.values()), {}
        node.target = replacement
        n += 1
    if n:
        graph.lint()
        if graph.owning_module is not None:
            graph.owning_module.recompile()
    logging.debug("Pre grad pass %s replaced %d sub-graphs", name, n)
    return n


def pre_grad_custom_pass(graph: torch.fx.Graph, traceable: bool = False) -> None:
    """Run the passes named in ``config.pre_grad_pass_options`` on ``graph`` in place (also the hook signature of Inductor's
    ``pre_grad_custom_pass`` config entry)."""
    for name in config.pre_grad_pass_options:
        if name not in PRE_GRAD_PASS_PATTERNS:
            raise AssertionError(f"Unknown pre_grad pass: {name}")
        pattern, replacement, traceable_replacement = PRE_GRAD_PASS_PATTERNS[name]
        done = run_pre_grad_pass(name, graph, pattern, traceable_replacement if traceable else replacement)
        counters[f"pre_grad_{name}"] = counters.get(f"pre_grad_{name}", 0) + done
