"""FusedAdamSWA: Adam on fp32 params + bf16 compute copy + stochastic-weight-average update in ONE persistent launch (csrc/mt_optim.cu AdamSwaOp).
Reference: apex/contrib/openfold_triton/fused_adam_swa.py:209-400 (one Triton multi-tensor kernel over pointer tables)."""
from __future__ import annotations

from enum import Enum, unique
from itertools import chain

import torch
from torch.optim import Optimizer

from ...ops import amp_C
from ...ops import reference as ref


@unique
class AdamMathType(Enum):
    ApexAdam = 0
    ApexAdamW = 1
    PyTorchAdam = 2


class FusedAdamSWA(Optimizer):
    def __init__(self, params, compute_params, swa_params, swa_decay_rate, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8,
                 adam_math_mode=AdamMathType.PyTorchAdam, weight_decay=0.0, amsgrad=False, set_grad_none=True, capturable=False,
                 master_weights=False):
        params, compute_params, swa_params = list(params), list(compute_params), list(swa_params)
        if not compute_params or not swa_params:
            raise ValueError("FusedAdamSWA requires both BF16 and SWA parameters.")
        if not len(params) == len(compute_params) == len(swa_params):
            raise ValueError("FusedAdamSWA expects params, bf16_params, and swa_params to have same length")
        if not all(p.shape == b.shape == s.shape for p, b, s in zip(params, compute_params, swa_params)):
            raise ValueError("FusedAdamSWA expects each state in params, bf16_params, abd swa_params to have same shape")
        if not all(p.is_contiguous() for p in chain(params, compute_params, swa_params)):
            raise ValueError("FusedAdamSWA expects all input params to be contiguous")
        if amsgrad or capturable or master_weights:
            raise NotImplementedError("amsgrad / capturable / master_weights are not supported by FusedAdamSWA")
        if not isinstance(adam_math_mode, AdamMathType):
            raise ValueError(f"Unknown Adam math mode {adam_math_mode}")
        super().__init__(params, dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay))
        self.adam_math_mode, self.set_grad_none = adam_math_mode, set_grad_none
        self.compute_param_groups = [{"params": compute_params}]
        self.swa_param_groups = [{"params": swa_params, "n_averaged": 0}]
        self.swa_decay_rate = swa_decay_rate
        self._tables = None

    @classmethod
    def from_optim(cls, adam_optimizer, fp32_params, bf16_params, swa_params, swa_decay_rate):
        """Take over from a (possibly checkpoint-restored) ``torch.optim.Adam`` with one param group: hyper-parameters, moments and the
        common step count carry over (reference fused_adam_swa.py:459-497)."""
        assert len(adam_optimizer.param_groups) == 1
        g = adam_optimizer.param_groups[0]
        opt = cls(params=fp32_params, compute_params=bf16_params, swa_params=swa_params, swa_decay_rate=swa_decay_rate, lr=g["lr"],
                  betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"], amsgrad=g.get("amsgrad", False),
                  adam_math_mode=AdamMathType.PyTorchAdam)
        sd = adam_optimizer.state_dict()
        steps = [v["step"] for v in sd["state"].values() if "step" in v]
        if steps and not all(float(s) == float(steps[0]) for s in steps):
            raise ValueError("FusedAdamSWA requires all parameters were updated by same steps!")
        sd["param_groups"][0].setdefault("bias_correction", True)
        sd["param_groups"][0]["step"] = int(float(steps[0])) if steps else 0
        for v in sd["state"].values():      # per-parameter step counters become the one group-level counter
            v.pop("step", None)
        opt.load_state_dict(sd)
        return opt

    @torch.no_grad()
    def step(self, closure=None, grad_clip_scale=None):
        """``grad_clip_scale`` (float or 0-dim tensor): factor applied to every gradient before the update — the caller's global-norm
        clipping coefficient (reference fused_adam_swa.py:372-380)."""
        if len(self.param_groups) != 1:
            raise RuntimeError("FusedAdamSWA does not support multiple param groups")
        loss = closure() if closure is not None else None
        group = self.param_groups[0]
        params, cparams, sparams = group["params"], self.compute_param_groups[0]["params"], self.swa_param_groups[0]["params"]
        group["step"] = group.get("step", 0) + 1
        for p in params:
            st = self.state[p]
            if not st:
                st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32), torch.zeros_like(p, dtype=torch.float32)
        # gradients come from the bf16 compute copies (OpenFold runs fwd/bwd on them)
        grads = [(c.grad if c.grad is not None else p.grad) for p, c in zip(params, cparams)]
        exp_avg, exp_avg_sq = [self.state[p]["exp_avg"] for p in params], [self.state[p]["exp_avg_sq"] for p in params]
        beta1, beta2 = group["betas"]
        mode = 1 if self.adam_math_mode == AdamMathType.ApexAdamW else 0
        # SWA: the first step copies, later steps swa = decay * swa + (1 - decay) * p
        n = self.swa_param_groups[0]["n_averaged"]
        a, b = (0.0, 1.0) if n == 0 else (self.swa_decay_rate, 1.0 - self.swa_decay_rate)
        cuda = params[0].is_cuda
        if cuda and all(p.dtype == torch.float32 for p in params) and all(s_.dtype == torch.float32 for s_ in sparams):
            # ONE persistent launch: Adam + SWA + compute copy (+ the clip factor read from the device), csrc/mt_optim.cu AdamSwaOp
            clip = None
            if grad_clip_scale is not None:
                clip = (grad_clip_scale if torch.is_tensor(grad_clip_scale) else torch.tensor(float(grad_clip_scale)))
                clip = clip.to(params[0].device, torch.float32).reshape(1)
            amp_C.multi_tensor_adam_swa(65536, [grads, params, exp_avg, exp_avg_sq, sparams, cparams], group["lr"], beta1, beta2, group["eps"],
                                        group["step"], mode, int(self.adam_math_mode == AdamMathType.PyTorchAdam), int(group["bias_correction"]),
                                        group["weight_decay"], a, b, clip)
        else:
            if grad_clip_scale is not None:
                grads = [g * grad_clip_scale for g in grads]
            lists = [grads, params, exp_avg, exp_avg_sq]
            if cuda:
                amp_C.multi_tensor_adam(65536, None, lists, group["lr"], beta1, beta2, group["eps"], group["step"], mode,
                                        int(group["bias_correction"]), group["weight_decay"])
            else:
                ref.multi_tensor_adam(lists, group["lr"], beta1, beta2, group["eps"], group["step"], mode, int(group["bias_correction"]),
                                      group["weight_decay"])
            noop = torch.zeros(1, dtype=torch.int32, device=params[0].device)
            amp_C.multi_tensor_axpby(65536, noop, [sparams, params, sparams], a, b, -1)
            amp_C.multi_tensor_scale(65536, noop, [params, cparams], 1.0)
        self.swa_param_groups[0]["n_averaged"] = n + 1
        if self.set_grad_none:
            for c in cparams:
                c.grad = None
        return loss
