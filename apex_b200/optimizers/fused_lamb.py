"""FusedLAMB and FusedMixedPrecisionLamb.

Reference: apex/optimizers/fused_lamb.py:5-244 and fused_mixed_precision_lamb.py:9-291. Behavioural parity: global grad-norm
clipping (``max_grad_norm``), per-tensor trust ratio (``use_nvlamb`` or weight_decay != 0), ``grad_averaging``, ``adam_w_mode``,
state in the parameter dtype (FusedLAMB) / fp32 masters created lazily on the first step (mixed precision).
The step is two persistent launches per bucket (update+norms, apply) instead of the reference's four.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..ops import amp_C
from ..ops import reference as ref
from ._base import BucketCache, adopt_foreign_state, partition_by_dtype, restore_fp32_state


def _global_grad_norm(tables, cpu_grads, device):
    """sqrt(sum of squared norms) over cached tables (slot 0 = grads) and loose CPU grads."""
    sq = []
    for tb in tables:
        n, _ = amp_C._norm(0, None, tb, False)
        sq.append(n.float() ** 2)
    for g in cpu_grads:
        sq.append((g.float() ** 2).sum().reshape(1))
    if not sq:
        return torch.zeros(1, device=device)
    return torch.stack([s.reshape(()) for s in sq]).sum().sqrt().reshape(1)


class FusedLAMB(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, amsgrad=False,
                 adam_w_mode=True, grad_averaging=True, set_grad_none=True, max_grad_norm=1.0, use_nvlamb=False):
        if amsgrad:
            raise RuntimeError("FusedLAMB does not support the AMSGrad variant.")
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                        grad_averaging=grad_averaging, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self.adam_w_mode = 1 if adam_w_mode else 0
        self.set_grad_none = set_grad_none
        self.use_nvlamb = use_nvlamb
        self._cache = BucketCache()
        self._parts: dict = {}

    def add_param_group(self, g):
        super().add_param_group(g)
        if hasattr(self, "_cache"):
            self._cache.clear()
            self._parts.clear()

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        adopt_foreign_state(self)
        self._cache.clear()

    def zero_grad(self, set_to_none: bool | None = None):
        if self.set_grad_none if set_to_none is None else set_to_none:
            for group in self.param_groups:
                for p in group["params"]:
                    p.grad = None
        else:
            super().zero_grad(set_to_none=False)

    def _bucket(self, gi, dtype, cands):
        """-> (table or None, cpu_lists or None)"""
        is_cuda = cands[0].is_cuda
        key = (gi, dtype)
        tb = self._cache.cached(key) if is_cuda else None
        if tb is not None:
            return tb, None
        members = [p for p in cands if p.grad is not None]
        if not members:
            return None, None
        for p in members:
            if p.grad.is_sparse:
                raise RuntimeError("FusedLAMB does not support sparse gradients, please consider SparseAdam instead")
            st = self.state[p]
            if len(st) == 0:
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
        lists = [[p.grad for p in members], list(members), [self.state[p]["exp_avg"] for p in members],
                 [self.state[p]["exp_avg_sq"] for p in members]]
        if not is_cuda:
            return None, lists
        if not _lib.available():
            raise _lib.gpu_required_error("FusedLAMB")
        return self._cache.build(key, cands, members, lists), None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        device = self.param_groups[0]["params"][0].device
        work = []
        for gi, group in enumerate(self.param_groups):
            parts = self._parts.get(gi)
            if parts is None:
                parts = self._parts[gi] = partition_by_dtype(group["params"])
            for dtype, cands in parts.items():
                if dtype not in (torch.float16, torch.bfloat16, torch.float32):
                    raise RuntimeError("FusedLAMB only support fp16, bf16 and fp32.")
                tb, cpu = self._bucket(gi, dtype, cands)
                if tb is not None or cpu is not None:
                    work.append((group, tb, cpu))
        ggn = _global_grad_norm([tb for _, tb, _ in work if tb is not None],
                                [g for _, _, cpu in work if cpu is not None for g in cpu[0]], device)
        max_grad_norm = self.defaults["max_grad_norm"]
        stepped = set()
        for group, tb, cpu in work:
            if id(group) not in stepped:
                group["step"] = group.get("step", 0) + 1
                stepped.add(id(group))
            beta1, beta2 = group["betas"]
            args = (group["lr"], beta1, beta2, group["eps"], group["step"], 1 if group["bias_correction"] else 0,
                    group["weight_decay"], 1 if group["grad_averaging"] else 0, self.adam_w_mode, ggn, max_grad_norm, self.use_nvlamb)
            if tb is not None:
                amp_C.multi_tensor_lamb(0, None, tb, *args)
            else:
                ref.multi_tensor_lamb(cpu, *args)
        return loss


class FusedMixedPrecisionLamb(torch.optim.Optimizer):
    """LAMB with fp32 master weights for reduced-precision model params; lr/step live on the device (graph-friendly)."""

    def __init__(self, params, lr=1e-3, step=0, bias_correction=True, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, amsgrad=False,
                 adam_w_mode=True, grad_averaging=True, max_grad_norm=1.0, use_nvlamb=False, reduced_precision_dtype=None):
        if amsgrad:
            raise RuntimeError("FusedLAMB does not support the AMSGrad variant.")
        defaults = dict(lr=torch.tensor(lr, dtype=torch.float32), step=torch.tensor([step], dtype=torch.int),
                        bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                        grad_averaging=grad_averaging, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        device = self.param_groups[0]["params"][0].device
        for group in self.param_groups:
            for k in ("lr", "step"):
                group[k] = group[k].to(device=device) if torch.is_tensor(group[k]) else torch.tensor(group[k], device=device)
        self.param_groups_full_precision: list = []
        self._step_supports_amp_scaling = True
        self.adam_w_mode = 1 if adam_w_mode else 0
        self.use_nvlamb = use_nvlamb
        self.reduced_precision_dtype = reduced_precision_dtype
        self._dummy_overflow_buf = torch.zeros(1, dtype=torch.int, device=device)
        self._cache = BucketCache()

    def state_dict(self):
        """torch's optimizer state plus the fp32 master parameters under ``"master_params"`` once they exist (the reference re-derives them
        from the rounded model weights on resume, which is lossy)."""
        sd = super().state_dict()
        if self.param_groups_full_precision:
            sd["master_params"] = [[None if m is None else m.detach().clone() for m in g["params"]] for g in self.param_groups_full_precision]
        return sd

    def load_state_dict(self, state_dict):
        # lr/step are tensors: keep them as device tensors after a reload (reference :73-138)
        super().load_state_dict(state_dict)
        masters = state_dict.get("master_params")
        if masters is not None:
            if len(self.param_groups_full_precision) == 0:
                self._setup_full_precision_params()
            with torch.no_grad():
                for g, saved in zip(self.param_groups_full_precision, masters):
                    for m, s_ in zip(g["params"], saved):
                        if m is not None and s_ is not None:
                            m.copy_(s_)
        device = self.param_groups[0]["params"][0].device
        for group in self.param_groups:
            for k, dt in (("lr", torch.float32), ("step", torch.int)):
                v = group[k]
                group[k] = (v if torch.is_tensor(v) else torch.tensor(v)).to(device=device, dtype=dt).reshape(-1)[:1].clone() \
                    if k == "step" else (v if torch.is_tensor(v) else torch.tensor(v)).to(device=device, dtype=dt)
        if self.reduced_precision_dtype is not None:
            restore_fp32_state(self, state_dict)
        for st in self.state.values():
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st and self.reduced_precision_dtype is not None and st[k].dtype != torch.float32:
                    st[k] = st[k].float()
        self._cache.clear()

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if hasattr(self, "_cache"):
            device = self.param_groups[0]["params"][0].device
            g = self.param_groups[-1]
            for k, dt in (("lr", torch.float32), ("step", torch.int)):
                v = g[k]
                g[k] = (v.clone() if torch.is_tensor(v) else torch.tensor(v)).to(device=device, dtype=dt)
            self._cache.clear()
            self.param_groups_full_precision = []

    def _setup_full_precision_params(self):
        # Done at the first step, not in __init__, so a DDP parameter broadcast that happens after construction is honoured.
        for pg in self.param_groups:
            self.param_groups_full_precision.append({"params": [
                p.clone().detach().to(dtype=torch.float32)
                if (self.reduced_precision_dtype is not None and p.dtype == self.reduced_precision_dtype) else None
                for p in pg["params"]]})

    @torch.no_grad()
    def step(self, closure=None, grad_scaler=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if len(self.param_groups_full_precision) == 0:
            self._setup_full_precision_params()
        device = self.param_groups[0]["params"][0].device
        grad_list = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        for g in self.param_groups:
            for p in g["params"]:
                assert g["params"][0].dtype == p.dtype, "Error: Parameters are not of the identical type: {} != {}".format(
                    g["params"][0].dtype, p.dtype)
        found_inf = (grad_scaler._check_inf_per_device(self)[device] if grad_scaler is not None else torch.zeros((1,), device=device))
        self._dummy_overflow_buf.copy_(found_inf)
        if grad_scaler:
            scale = grad_scaler._get_scale_async()
            inv_scale = scale.double().reciprocal().float()
        else:
            scale = torch.ones((1,), device=device)
            inv_scale = torch.ones((1,), device=device)
        max_grad_norm = self.defaults["max_grad_norm"] * scale  # the norm is taken on still-scaled grads
        grad_norm = amp_C.multi_tensor_l2norm(65536, self._dummy_overflow_buf, [grad_list], False)[0]

        for gi, (group, group_full) in enumerate(zip(self.param_groups, self.param_groups_full_precision)):
            beta1, beta2 = group["betas"]
            group["step"] += (self._dummy_overflow_buf != 1).to(torch.int)
            members = [(p, pf) for p, pf in zip(group["params"], group_full["params"]) if p.grad is not None]
            if not members:
                continue
            is_cuda = members[0][0].is_cuda
            tb = self._cache.cached(gi) if is_cuda else None
            if tb is None:
                lists = [[], [], [], []]
                use_master = self.reduced_precision_dtype is not None and members[0][1] is not None
                if use_master:
                    lists.append([])
                for p, pf in members:
                    st = self.state[p]
                    if len(st) == 0:
                        dt = torch.float32 if use_master else p.dtype
                        st["exp_avg"] = torch.zeros_like(p, dtype=dt)
                        st["exp_avg_sq"] = torch.zeros_like(p, dtype=dt)
                    lists[0].append(p.grad)
                    lists[1].append(pf if use_master else p)
                    lists[2].append(st["exp_avg"])
                    lists[3].append(st["exp_avg_sq"])
                    if use_master:
                        lists[4].append(p)
                if is_cuda:
                    if not _lib.available():
                        raise _lib.gpu_required_error("FusedMixedPrecisionLamb")
                    ps = [p for p, _ in members]
                    tb = self._cache.build(gi, list(group["params"]), ps, lists, grad_slot=0, param_slot=(4 if use_master else 1))
            args = (group["lr"], beta1, beta2, group["eps"], group["step"], 1 if group["bias_correction"] else 0, group["weight_decay"],
                    1 if group["grad_averaging"] else 0, self.adam_w_mode, grad_norm, max_grad_norm, self.use_nvlamb, found_inf, inv_scale)
            if tb is not None:
                amp_C.multi_tensor_lamb_mp(0, self._dummy_overflow_buf, tb, *args)
            else:
                ref.multi_tensor_lamb_mp(self._dummy_overflow_buf, lists, *args)
        return loss
