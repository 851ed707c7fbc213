"""FusedSGD, FusedNovoGrad, FusedAdagrad on the cached device-table engine.

Reference: apex/optimizers/fused_sgd.py:8-284, fused_novograd.py:5-255, fused_adagrad.py:5-134.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..ops import amp_C
from ..ops import reference as ref
from ._base import CHUNK, BucketCache, adopt_foreign_state, partition_by_dtype


class _TableOptimizer(torch.optim.Optimizer):
    set_grad_none = True

    def _init_cache(self):
        self._cache = BucketCache()
        self._parts: dict = {}

    def add_param_group(self, g):
        super().add_param_group(g)
        if hasattr(self, "_cache"):
            self._cache.clear()
            self._parts.clear()

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        adopt_foreign_state(self)   # checkpoints of the torch.optim counterpart: missing hyper-parameters, per-parameter step counters
        self._cache.clear()

    def zero_grad(self, set_to_none: bool | None = None):
        if self.set_grad_none if set_to_none is None else set_to_none:
            for group in self.param_groups:
                for p in group["params"]:
                    p.grad = None
        else:
            super().zero_grad(set_to_none=False)

    def _partition(self, gi, group):
        parts = self._parts.get(gi)
        if parts is None:
            parts = self._parts[gi] = partition_by_dtype(group["params"])
        return parts


class FusedSGD(_TableOptimizer):
    """SGD with momentum / dampening / nesterov / ``wd_after_momentum``; momentum buffers are created on the first step and that
    step runs with ``first_run`` semantics (buffer := grad). ``scale`` of the most recent loss scale can be folded into the
    update via ``most_recent_scale`` (reference :276)."""

    def __init__(self, params, lr, momentum=0, dampening=0, weight_decay=0, nesterov=False, wd_after_momentum=False,
                 materialize_master_grads=True, set_grad_none=False):
        if lr is None or lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if momentum < 0.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(params, defaults)
        self.wd_after_momentum = wd_after_momentum
        self.materialize_master_grads = materialize_master_grads
        self.most_recent_scale = 1.0
        self.scale_set_by_backward = False
        self.set_grad_none = set_grad_none
        self._init_cache()

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("nesterov", False)

    def get_momentums(self, params):
        """(momentum buffers of ``params``, first_run): buffers are created on demand, and ``first_run`` is reported by the last parameter
        examined, as in the reference (fused_sgd.py:137-152)."""
        momentums, first_run = [], True
        for p in params:
            st = self.state[p]
            first_run = "momentum_buffer" not in st
            if first_run:
                st["momentum_buffer"] = torch.zeros_like(p.data)
            momentums.append(st["momentum_buffer"])
        return momentums, first_run

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            for dtype, cands in self._partition(gi, group).items():
                is_cuda = cands[0].is_cuda
                key = (gi, dtype)
                tb = self._cache.cached(key) if is_cuda else None
                launches = []   # (table or tensor lists, first_run)
                if tb is not None:
                    launches.append((tb, False))
                else:
                    members = [p for p in cands if p.grad is not None]
                    if not members:
                        continue
                    if any(p.grad.is_sparse for p in members):
                        raise RuntimeError("FusedSGD does not support sparse gradients")
                    # A momentum buffer starts as the first gradient the parameter ever sees (torch.optim.SGD semantics). ``first_run`` is
                    # a per-launch flag, so parameters that receive their first gradient later than the others get their own launch for
                    # that one step (the reference applies the flag of the last parameter to the whole list, fused_sgd.py:137-152).
                    fresh = [p for p in members if "momentum_buffer" not in self.state[p]]
                    for p in fresh:
                        self.state[p]["momentum_buffer"] = torch.zeros_like(p)
                    seen = [p for p in members if not any(p is q for q in fresh)] if fresh else members

                    def lists_of(ps):
                        return [[p.grad for p in ps], list(ps), [self.state[p]["momentum_buffer"] for p in ps]]

                    if is_cuda and not _lib.available():
                        raise _lib.gpu_required_error("FusedSGD")
                    if fresh and seen:
                        launches += [(lists_of(seen), False), (lists_of(fresh), True)]   # one-off; the table is cached from the next step on
                    else:
                        lists = lists_of(members)
                        launches.append((self._cache.build(key, cands, members, lists) if is_cuda else lists, bool(fresh)))
                for target, first_run in launches:
                    args = (group["weight_decay"], group["momentum"], group["dampening"], group["lr"], group["nesterov"], first_run,
                            self.wd_after_momentum, 1.0 / self.most_recent_scale)
                    if is_cuda:
                        amp_C.multi_tensor_sgd(CHUNK, None, target, *args)   # CHUNK matters for the one-off list launches (tables carry their own)
                    else:
                        ref.multi_tensor_sgd(None, target, *args)
        self.most_recent_scale = 1.0
        self.scale_set_by_backward = False
        return loss


class FusedNovoGrad(_TableOptimizer):
    """NovoGrad: per-TENSOR second moment (a scalar norm per parameter), kept as one float tensor per (group, dtype)."""

    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False,
                 reg_inside_moment=False, grad_averaging=True, norm_type=2, init_zero=False, set_grad_none=True):
        if amsgrad:
            raise RuntimeError("FusedNovoGrad does not support the AMSGrad variant.")
        if norm_type not in (0, 2):
            raise RuntimeError("FusedNovoGrad only support l2/inf norm now.")
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                        grad_averaging=grad_averaging, norm_type=norm_type, init_zero=init_zero)
        super().__init__(params, defaults)
        self.moment_mode = 0 if reg_inside_moment else 1
        self.set_grad_none = set_grad_none
        self._init_cache()
        self._norms: dict = {}

    def load_state_dict(self, state_dict):
        """Restores the per-tensor second moments too. They live in ``group["exp_avg_sq"]``: here a dict {dtype name: norms of that dtype's
        parameters, in group order}; the reference's two-element list [16-bit norms, fp32 norms] (fused_novograd.py:180-218) is accepted."""
        super().load_state_dict(state_dict)
        self._norms.clear()
        names = {str(d): d for d in (torch.float16, torch.bfloat16, torch.float32, torch.float64)}
        for gi, group in enumerate(self.param_groups):
            saved = group.get("exp_avg_sq")
            if not saved or not group["params"]:
                continue
            dev = group["params"][0].device
            present = {p.dtype for p in group["params"]}
            if isinstance(saved, (list, tuple)):
                half = next((d for d in (torch.float16, torch.bfloat16) if d in present), torch.float16)
                saved = {str(d): t for d, t in zip((half, torch.float32), saved) if torch.is_tensor(t) and t.numel()}
            restored = {}
            for name, t in saved.items():
                if name in names and torch.is_tensor(t):
                    restored[name] = t.to(device=dev, dtype=torch.float32)
                    self._norms[(gi, names[name])] = restored[name]   # the SAME tensor the kernel updates in place
            group["exp_avg_sq"] = restored

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            beta1, beta2 = group["betas"]
            group["step"] = group.get("step", 0) + 1
            for dtype, cands in self._partition(gi, group).items():
                is_cuda = cands[0].is_cuda
                key = (gi, dtype)
                tb = self._cache.cached(key) if is_cuda else None
                lists = None
                if tb is None or key not in self._norms:
                    members = [p for p in cands if p.grad is not None]
                    if not members:
                        continue
                    for p in members:
                        if "exp_avg" not in self.state[p]:
                            self.state[p]["exp_avg"] = torch.zeros_like(p)
                    lists = [[p.grad for p in members], list(members), [self.state[p]["exp_avg"] for p in members]]
                    norms = self._norms.get(key)
                    if norms is None or norms.numel() != len(members):
                        if group["init_zero"]:
                            norms = torch.zeros(len(members), dtype=torch.float32, device=members[0].device)
                        elif group["norm_type"] == 0:
                            norms = torch.stack([p.grad.float().abs().max() for p in members])
                        else:
                            norms = torch.stack([p.grad.float().norm() for p in members])
                        self._norms[key] = norms
                        group.setdefault("exp_avg_sq", {})
                        if isinstance(group["exp_avg_sq"], dict):
                            group["exp_avg_sq"][str(dtype)] = norms
                    if is_cuda:
                        if not _lib.available():
                            raise _lib.gpu_required_error("FusedNovoGrad")
                        tb = self._cache.build(key, cands, members, lists)
                args = (self._norms[key], group["lr"], beta1, beta2, group["eps"], group["step"], 1 if group["bias_correction"] else 0,
                        group["weight_decay"], 1 if group["grad_averaging"] else 0, self.moment_mode, group["norm_type"])
                if tb is not None:
                    amp_C.multi_tensor_novograd(0, None, tb, *args)
                else:
                    ref.multi_tensor_novograd(lists, *args)
        return loss


class FusedAdagrad(_TableOptimizer):
    def __init__(self, params, lr=1e-2, eps=1e-10, weight_decay=0.0, set_grad_none=True, adagrad_w_mode=False):
        defaults = dict(lr=lr, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.adagrad_w_mode = 1 if adagrad_w_mode else 0
        self.set_grad_none = set_grad_none
        self._init_cache()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            for dtype, cands in self._partition(gi, group).items():
                is_cuda = cands[0].is_cuda
                key = (gi, dtype)
                tb = self._cache.cached(key) if is_cuda else None
                lists = None
                if tb is None:
                    members = [p for p in cands if p.grad is not None]
                    if not members:
                        continue
                    for p in members:
                        if p.grad.is_sparse:
                            raise RuntimeError("FusedAdagrad does not support sparse gradients")
                        if "sum" not in self.state[p]:
                            self.state[p]["sum"] = torch.zeros_like(p)
                    lists = [[p.grad for p in members], list(members), [self.state[p]["sum"] for p in members]]
                    if is_cuda:
                        if not _lib.available():
                            raise _lib.gpu_required_error("FusedAdagrad")
                        tb = self._cache.build(key, cands, members, lists)
                args = (group["lr"], group["eps"], self.adagrad_w_mode, group["weight_decay"])
                if tb is not None:
                    amp_C.multi_tensor_adagrad(0, None, tb, *args)
                else:
                    ref.multi_tensor_adagrad(lists, *args)
        return loss
