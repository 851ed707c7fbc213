"""FusedAdam — Adam/AdamW whose whole step is ONE persistent sm_100a launch per (param group, dtype).

API and semantics follow the reference ``apex.optimizers.FusedAdam`` (apex/optimizers/fused_adam.py:5-355): fp32 moments for
every parameter dtype, ``adam_w_mode``, ``capturable`` (lr/step on device, GradScaler found_inf/inv_scale consumed in-kernel,
``step += found_inf != 1``), ``master_weights`` (fp32 master copy, requires capturable), ``set_grad_none``.
Unlike the reference there is also a CPU path (plain PyTorch math) so the same script runs without a GPU.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..ops import amp_C
from ..ops import reference as ref
from ._base import BucketCache, adopt_foreign_state, flat_state_like, partition_by_dtype, restore_fp32_state


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, adam_w_mode=True, weight_decay=0.0,
                 amsgrad=False, set_grad_none=True, capturable=False, master_weights=False):
        if amsgrad:
            raise RuntimeError("FusedAdam does not support the AMSGrad variant.")
        if master_weights and not capturable:
            raise RuntimeError("Master weights is currently only supported with the capturable version.")
        lr = torch.tensor(lr, dtype=torch.float32) if capturable else lr
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.adam_w_mode = 1 if adam_w_mode else 0
        self.set_grad_none = set_grad_none
        self.capturable = capturable
        self.master_weights = master_weights
        self.param_groups_master = []
        for pg in self.param_groups:
            self.param_groups_master.append(
                {"params": [p.clone().detach().float() if master_weights else None for p in pg["params"]]})
        if capturable:
            for group in self.param_groups:
                if group["params"]:
                    group["lr"] = group["lr"].to(device=group["params"][0].device)
            self._step_supports_amp_scaling = True
        self._cache = BucketCache()
        self._parts: dict = {}
        self._dummy_overflow_buf = None

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if hasattr(self, "_cache"):
            self._cache.clear()
            self._parts.clear()
            if self.master_weights:
                self.param_groups_master.append({"params": [p.clone().detach().float() for p in self.param_groups[-1]["params"]]})
            else:
                self.param_groups_master.append({"params": [None for _ in self.param_groups[-1]["params"]]})

    def state_dict(self):
        """torch's optimizer state plus, with ``master_weights``, the fp32 master copies under ``"master_params"`` (the reference keeps them out
        of the checkpoint and re-derives them from the rounded model weights on resume, which is lossy)."""
        sd = super().state_dict()
        if self.master_weights:
            sd["master_params"] = [[None if m is None else m.detach().clone() for m in g["params"]] for g in self.param_groups_master]
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        adopt_foreign_state(self)   # e.g. a torch.optim.Adam(W) checkpoint
        if self.capturable:         # lr / step live on the device in capturable mode, whatever the checkpoint stored
            for group in self.param_groups:
                if not group["params"]:
                    continue
                dev = group["params"][0].device
                group["lr"] = torch.as_tensor(group["lr"], dtype=torch.float32).to(dev)
                if "step" in group:
                    group["step"] = torch.as_tensor(group["step"], dtype=torch.int).reshape(-1)[:1].to(dev).clone()
        # torch casts loaded state to the parameter dtype; the kernels keep fp32 moments for every parameter dtype.
        restore_fp32_state(self, state_dict)
        for st in self.state.values():
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st and st[k].dtype != torch.float32:
                    st[k] = st[k].float()
        masters = state_dict.get("master_params")
        if masters is not None and self.master_weights:
            with torch.no_grad():
                for g, saved in zip(self.param_groups_master, masters):
                    for m, s_ in zip(g["params"], saved):
                        if m is not None and s_ is not None:
                            m.copy_(s_)
        self._cache.clear()

    def zero_grad(self, set_to_none: bool | None = None):
        if self.set_grad_none if set_to_none is None else set_to_none:
            for group in self.param_groups:
                for p in group["params"]:
                    p.grad = None
        else:
            super().zero_grad(set_to_none=False)

    def _noop(self, device):
        if self._dummy_overflow_buf is None or self._dummy_overflow_buf.device != device:
            self._dummy_overflow_buf = torch.zeros(1, dtype=torch.int, device=device)
        return self._dummy_overflow_buf

    def _init_state(self, members):
        fresh = [p for p in members if len(self.state[p]) == 0]
        ms, vs = flat_state_like(fresh), flat_state_like(fresh)
        for p, m, v in zip(fresh, ms, vs):
            self.state[p]["exp_avg"] = m
            self.state[p]["exp_avg_sq"] = v

    @torch.no_grad()
    def step(self, closure=None, grads=None, output_params=None, scale=None, grad_norms=None, grad_scaler=None):
        if any(a is not None for a in (grads, output_params, scale, grad_norms)):
            raise RuntimeError("FusedAdam has been updated. Simply initialize it identically to torch.optim.Adam, and call step() "
                               "with no arguments.")
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()

        for gi, (group, group_master) in enumerate(zip(self.param_groups, self.param_groups_master)):
            if not group["params"]:
                continue
            device = group["params"][0].device
            bias_correction = 1 if group["bias_correction"] else 0
            beta1, beta2 = group["betas"]
            if "step" in group:
                if self.capturable:
                    noop = self._noop(device)
                    group["step"] += (noop != 1).to(torch.int)
                else:
                    group["step"] += 1
            else:
                group["step"] = 1 if not self.capturable else torch.tensor([1], dtype=torch.int, device=device)

            found_inf = inv_scale = None
            if self.capturable:
                noop = self._noop(device)
                if grad_scaler is not None:
                    found_inf = grad_scaler._check_inf_per_device(self)[device]
                    noop.copy_(found_inf)
                    scale_t = grad_scaler._get_scale_async()
                    inv_scale = scale_t.double().reciprocal().float()
                else:
                    inv_scale = torch.ones(1, device=device)

            parts = self._parts.get(gi)
            if parts is None:
                parts = self._parts[gi] = partition_by_dtype(group["params"])
                self._master_of = getattr(self, "_master_of", {})
                for p, pm in zip(group["params"], group_master["params"]):
                    self._master_of[id(p)] = pm
            for dtype, cands in parts.items():
                if dtype not in (torch.float16, torch.bfloat16, torch.float32):
                    raise RuntimeError("FusedAdam only support fp16, bf16 and fp32.")
                is_cuda = cands[0].is_cuda
                use_master = self.master_weights and dtype != torch.float32
                key = (gi, dtype)
                tb = self._cache.cached(key) if is_cuda else None
                if tb is None:
                    members = [p for p in cands if p.grad is not None]
                    if not members:
                        continue
                    if any(p.grad.is_sparse for p in members):
                        raise RuntimeError("FusedAdam does not support sparse gradients, please consider SparseAdam instead")
                    self._init_state(members)
                    lists = [[p.grad for p in members], list(members), [self.state[p]["exp_avg"] for p in members],
                             [self.state[p]["exp_avg_sq"] for p in members]]
                    if use_master:
                        lists.append([self._master_of[id(p)] for p in members])
                    if is_cuda:
                        if not _lib.available():
                            raise _lib.gpu_required_error("FusedAdam")
                        tb = self._cache.build(key, cands, members, lists)
                    else:
                        self._cpu_step(lists, group, beta1, beta2, bias_correction, inv_scale)
                        continue
                if self.capturable:
                    amp_C.multi_tensor_adam_capturable(0, self._noop(device), tb, group["lr"], beta1, beta2, group["eps"], group["step"],
                                                       self.adam_w_mode, bias_correction, group["weight_decay"], inv_scale)
                else:
                    amp_C.multi_tensor_adam(0, None, tb, group["lr"], beta1, beta2, group["eps"], group["step"], self.adam_w_mode,
                                            bias_correction, group["weight_decay"])
        return loss

    def _cpu_step(self, lists, group, beta1, beta2, bias_correction, inv_scale):
        if self.capturable:
            ref.multi_tensor_adam_capturable(self._dummy_overflow_buf, lists, group["lr"], beta1, beta2, group["eps"], group["step"],
                                             self.adam_w_mode, bias_correction, group["weight_decay"], inv_scale)
        else:
            ref.multi_tensor_adam(lists, group["lr"], beta1, beta2, group["eps"], group["step"], self.adam_w_mode, bias_correction,
                                  group["weight_decay"])
