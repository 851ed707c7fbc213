"""Shared machinery for the fused optimizers: per-(group, dtype) cached device tensor tables.

The reference optimizers rebuild four python lists over every parameter on every step and hand them to a launcher that
re-validates and re-packs them (apex/optimizers/fused_adam.py:190-232, csrc/multi_tensor_apply.cuh:35-102). Here the table
for a (param group, dtype) bucket is built once; each step a single C++ call re-reads the ``.grad`` pointers and uploads
that one column only if it changed. ``step()`` is O(1) python work per bucket.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..ops.amp_C import TensorTable

import os as _os

CHUNK = int(_os.environ.get("APEX_B200_MT_CHUNK", 65536))  # elements per work item (default: the granule of the reference's multi_tensor_applier)


class BucketCache:
    """One cached TensorTable per key (group index, dtype, ...)."""

    def __init__(self):
        self._tables: dict = {}

    def clear(self):
        self._tables.clear()

    def cached(self, key):
        """The cached table with its gradient column refreshed, or None if it must be (re)built."""
        tb = self._tables.get(key)
        if tb is None:
            return None
        return tb if tb.refresh_grads() else None

    def build(self, key, candidates, members, lists, grad_slot=0, param_slot=1, chunk=CHUNK):
        """candidates: every param that could belong to this bucket; members: those that currently have a grad;
        lists: the tensor lists (``.grad`` of members in ``grad_slot``)."""
        tb = TensorTable(lists, chunk)
        if grad_slot is not None:
            tb.track_grads(grad_slot, param_slot)
            mem = {id(p) for p in members}
            tb.track_universe(candidates, [id(p) in mem for p in candidates])
        self._tables[key] = tb
        return tb


def partition_by_dtype(params):
    """{dtype: [params]} for CUDA/CPU dense params, preserving order."""
    out: dict = {}
    for p in params:
        out.setdefault(p.dtype, []).append(p)
    return out


def flat_state_like(params, dtype=torch.float32):
    """One flat zero buffer + per-parameter views (fewer allocations, contiguous optimizer state)."""
    if not params:
        return []
    total = sum(p.numel() for p in params)
    flat = torch.zeros(total, dtype=dtype, device=params[0].device)
    out, off = [], 0
    for p in params:
        n = p.numel()
        if p.is_contiguous():
            out.append(flat[off:off + n].view(p.shape))
        else:
            out.append(torch.zeros_like(p, dtype=dtype))
        off += n
    return out


def use_native(params) -> bool:
    """CUDA params -> native kernels (hard error if the library is missing); CPU params -> PyTorch reference path."""
    any_cuda = any(p.is_cuda for p in params)
    if any_cuda and not _lib.available():
        raise _lib.gpu_required_error("fused optimizer")
    return any_cuda


def restore_fp32_state(optimizer, state_dict, keys=("exp_avg", "exp_avg_sq")) -> None:
    """``torch.optim.Optimizer.load_state_dict`` casts every floating-point state tensor to the dtype of its parameter; moments kept in fp32
    for 16-bit parameters would be rounded on every resume. Re-install them from the checkpoint at full precision."""
    by_index = {}
    for group, saved in zip(optimizer.param_groups, state_dict["param_groups"]):
        for p, idx in zip(group["params"], saved["params"]):
            by_index[idx] = p
    for idx, st in state_dict["state"].items():
        p = by_index.get(idx)
        if p is None:
            continue
        for k in keys:
            v = st.get(k)
            if torch.is_tensor(v) and v.is_floating_point():
                optimizer.state[p][k] = v.detach().to(device=p.device, dtype=torch.float32).clone()


def adopt_foreign_state(optimizer) -> None:
    """Make a checkpoint written by a ``torch.optim`` optimizer of the same family loadable: hyper-parameters it does not know
    (``bias_correction``, ``grad_averaging``, ...) come from this optimizer's defaults, and torch's per-parameter ``step`` counters become the
    per-group counter these optimizers keep (they must agree inside a group, as they do after ordinary training)."""
    for group in optimizer.param_groups:
        for k, v in optimizer.defaults.items():
            group.setdefault(k, v)
        steps = []
        for p in group["params"]:
            st = optimizer.state.get(p)
            if st is not None and "step" in st:
                steps.append(int(float(st.pop("step"))))
        if steps and not group.get("step"):
            if len(set(steps)) != 1:
                raise ValueError("cannot adopt a checkpoint whose parameters of one group were updated a different number of times")
            group["step"] = steps[0]
