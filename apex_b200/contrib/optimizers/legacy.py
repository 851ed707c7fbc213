"""Deprecated contrib optimizers kept for API parity. Reference: apex/contrib/optimizers/{fused_adam,fused_lamb,fused_sgd,fp16_optimizer}.py
(967 lines) over ``fused_adam_cuda`` (single-tensor adam with explicit grads / output_params / scale, reversible adam, strided
check_finite, maybe_cast_mt). The explicit-argument ``step(grads=, output_params=, scale=, grad_norms=)`` contract is honoured on top
of the multi-tensor engine: gradients are unscaled (and norm-clipped) by the scale kernel, the update is the FusedAdam kernel, reduced
precision copies are written by the multi-tensor cast."""
from __future__ import annotations

import types

import torch

from ...ops import amp_C
from ...optimizers import FusedLAMB as _FusedLAMB
from ...optimizers import FusedSGD as _FusedSGD
from ...optimizers._base import CHUNK


class FusedAdam(torch.optim.Optimizer):
    """Legacy contrib FusedAdam: ``step(grads=None, output_params=None, scale=1.0, grad_norms=None)``; ``eps_inside_sqrt``;
    ``max_grad_norm`` clipping folded into the gradient scale."""

    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, eps_inside_sqrt=False, weight_decay=0.0,
                 max_grad_norm=0.0, amsgrad=False, use_mt=False, amp_scale_adjustment=1.0):
        if amsgrad:
            raise RuntimeError("FusedAdam does not support the AMSGrad variant.")
        super().__init__(params, dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                                      max_grad_norm=max_grad_norm))
        self.eps_mode = 0 if eps_inside_sqrt else 1
        self._amp_scale_adjustment = amp_scale_adjustment

    @torch.no_grad()
    def step(self, closure=None, grads=None, output_params=None, scale=1.0, grad_norms=None):
        loss = closure() if closure is not None else None
        if hasattr(self, "_amp_stash"):
            grads, output_params = self._amp_stash.grads, self._amp_stash.output_params
            scale, grad_norms = self._amp_stash.scale * self._amp_scale_adjustment, self._amp_stash.grad_norms

        def _group(x):
            if x is None:
                return [None] * len(self.param_groups)
            if isinstance(x, types.GeneratorType):
                return [list(x)]
            return [x] if not isinstance(x[0], list) else x

        grads_group, out_group = _group(grads), _group(output_params)
        grad_norms = grad_norms if grad_norms is not None else [None] * len(self.param_groups)
        for group, gs, outs, gnorm in zip(self.param_groups, grads_group, out_group, grad_norms):
            gs = gs if gs is not None else [None] * len(group["params"])
            outs = outs if outs is not None else [None] * len(group["params"])
            combined_scale = scale
            if group["max_grad_norm"] > 0 and gnorm is not None:
                clip = ((float(gnorm) / scale) + 1e-6) / group["max_grad_norm"]
                if clip > 1:
                    combined_scale = clip * scale
            group["step"] = group.get("step", 0) + 1
            g_l, p_l, m_l, v_l, o_src, o_dst = [], [], [], [], [], []
            for p, g, o in zip(group["params"], gs, outs):
                g = g if g is not None else p.grad
                if g is None:
                    continue
                st = self.state[p]
                if not st:
                    st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32), torch.zeros_like(p, dtype=torch.float32)
                gf = g.float() / combined_scale if combined_scale != 1.0 or g.dtype != p.dtype else g
                g_l.append(gf.to(p.dtype) if gf.dtype != p.dtype and p.dtype != torch.float32 else gf.to(torch.float32) if p.dtype == torch.float32 else gf)
                p_l.append(p)
                m_l.append(st["exp_avg"])
                v_l.append(st["exp_avg_sq"])
                if o is not None:
                    o_src.append(p)
                    o_dst.append(o)
            if not p_l:
                continue
            beta1, beta2 = group["betas"]
            amp_C.multi_tensor_adam(65536, None, [g_l, p_l, m_l, v_l], group["lr"], beta1, beta2, group["eps"], group["step"], 0,
                                    int(group["bias_correction"]), group["weight_decay"])
            if o_dst:
                amp_C.multi_tensor_cast(65536, None, [o_src, o_dst])
        return loss


class FusedLAMB(_FusedLAMB):
    """Legacy contrib FusedLAMB (same math as apex_b200.optimizers.FusedLAMB)."""


class FusedSGD(_FusedSGD):
    """Legacy contrib FusedSGD (reference contrib/optimizers/fused_sgd.py:129-260): ``step(grads=, output_params=, scale=)`` where
    ``group['params']`` are the fp32 masters, ``output_params`` the model weights (fp16 ones receive a copy of the update inside the
    same kernel launch) and ``scale`` divides the gradients. Without ``grads`` it behaves like ``apex_b200.optimizers.FusedSGD``
    (the reference raises there; this is a superset)."""

    @torch.no_grad()
    def step(self, closure=None, grads=None, output_params=None, scale=1.0, grad_norms=None):
        if hasattr(self, "_amp_stash"):
            raise RuntimeError("apex.contrib.optimizers.FusedSGD should not be used with AMP.")
        if grads is None and output_params is None:
            return super().step(closure)
        if grads is None or output_params is None:
            raise RuntimeError("apex.contrib.optimizers.FusedSGD needs both grads and output_params (FP16_Optimizer provides them).")
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()

        def _per_group(x):
            x = list(x) if isinstance(x, types.GeneratorType) else x
            return x if isinstance(x[0], list) else [x]

        for group, gs, outs in zip(self.param_groups, _per_group(grads), _per_group(output_params)):
            if gs is None or outs is None:
                raise RuntimeError("apex.contrib.optimizers.FusedSGD only works when all parameters require grad.")
            # two launches per group: model weights in fp32 (3 lists) and in half precision (4 lists: the kernel also writes the copy)
            for half in (True, False):
                sel = [(g, m, o) for g, m, o in zip(gs, group["params"], outs) if (o.dtype != torch.float32) == half]
                if not sel:
                    continue
                masters = [m for _, m, _ in sel]
                moms, first_run = self.get_momentums(masters)
                lists = [[g for g, _, _ in sel], masters, moms] + ([[o for _, _, o in sel]] if half else [])
                amp_C.multi_tensor_sgd(CHUNK, self._noop(masters[0]), lists, group["weight_decay"], group["momentum"], group["dampening"],
                                       group["lr"], group["nesterov"], first_run, self.wd_after_momentum, 1.0 / scale)
        return loss

    def _noop(self, like):
        buf = getattr(self, "_legacy_noop", None)
        if buf is None or buf.device != like.device:
            buf = self._legacy_noop = torch.zeros(1, dtype=torch.int32, device=like.device)
        return buf


class FP16_Optimizer:
    """Cut-down static / dynamic loss-scale wrapper around a fused optimizer (reference fp16_optimizer.py:5-248): fp16 model
    params, fp32 master copies, ``backward(loss)`` scales the loss, ``step()`` unscales, checks overflow, updates, copies back."""

    def __init__(self, init_optimizer, static_loss_scale=1.0, dynamic_loss_scale=False, dynamic_loss_args=None, verbose=True):
        self.optimizer = init_optimizer
        self.fp16_groups, self.fp32_groups = [], []
        for group in self.optimizer.param_groups:
            fp16 = list(group["params"])
            fp32 = [p.detach().clone().float().requires_grad_(True) for p in fp16]
            self.fp16_groups.append(fp16)
            self.fp32_groups.append(fp32)
            group["params"] = fp32
        if dynamic_loss_scale:
            self.dynamic_loss_scale = True
            a = dynamic_loss_args or {}
            self.cur_scale = a.get("init_scale", 2 ** 16)
            self.scale_factor, self.scale_window = a.get("scale_factor", 2), a.get("scale_window", 1000)
            self.cur_iter, self.last_overflow_iter = 0, -1
        else:
            self.dynamic_loss_scale, self.cur_iter, self.cur_scale = False, 0, static_loss_scale
        self.verbose = verbose

    def zero_grad(self, set_grads_to_None=True):
        for group in self.fp16_groups:
            for p in group:
                if set_grads_to_None:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.detach_().zero_()

    def backward(self, loss):
        (loss.float() * self.cur_scale).backward()

    @torch.no_grad()
    def step(self, closure=None):
        grads = [p.grad for g in self.fp16_groups for p in g if p.grad is not None]
        if not grads:
            return
        noop = torch.zeros(1, dtype=torch.int32, device=grads[0].device)
        norm, _ = amp_C.multi_tensor_l2norm(65536, noop, [grads], False)
        overflow = not bool(torch.isfinite(norm))
        self._update_scale(overflow)
        if overflow:
            if self.verbose:
                print(f"[FP16_Optimizer] OVERFLOW! Skipping step. Attempted loss scale: {self.cur_scale}")
            return
        for fp16, fp32 in zip(self.fp16_groups, self.fp32_groups):
            for p, m in zip(fp16, fp32):
                m.grad = None if p.grad is None else p.grad.float() / self.cur_scale_prev
        self.optimizer.step()
        for fp16, fp32 in zip(self.fp16_groups, self.fp32_groups):
            for p, m in zip(fp16, fp32):
                p.copy_(m)

    def _update_scale(self, skip):
        self.cur_scale_prev = self.cur_scale
        if self.dynamic_loss_scale:
            if skip:
                self.cur_scale = max(self.cur_scale / self.scale_factor, 1)
                self.last_overflow_iter = self.cur_iter
            elif (self.cur_iter - self.last_overflow_iter) % self.scale_window == 0:
                self.cur_scale *= self.scale_factor
        self.cur_iter += 1

    # promoted so that ``fp16_optimizer.state`` / ``.param_groups`` can be read and assigned (e.g. to adjust the learning rate)
    state = property(lambda self: self.optimizer.state, lambda self, value: setattr(self.optimizer, "state", value))
    param_groups = property(lambda self: self.optimizer.param_groups, lambda self, value: setattr(self.optimizer, "param_groups", value))

    def state_dict(self):
        sd = {"dynamic_loss_scale": self.dynamic_loss_scale, "cur_scale": self.cur_scale, "cur_iter": self.cur_iter,
              "optimizer_state_dict": self.optimizer.state_dict(), "fp32_groups": self.fp32_groups}
        if self.dynamic_loss_scale:
            sd.update(last_overflow_iter=self.last_overflow_iter, scale_factor=self.scale_factor, scale_window=self.scale_window)
        return sd

    def load_state_dict(self, state_dict):
        sd = state_dict
        self.dynamic_loss_scale, self.cur_scale, self.cur_iter = sd["dynamic_loss_scale"], sd["cur_scale"], sd["cur_iter"]
        if self.dynamic_loss_scale:
            self.last_overflow_iter, self.scale_factor, self.scale_window = sd["last_overflow_iter"], sd["scale_factor"], sd["scale_window"]
        self.optimizer.load_state_dict(sd["optimizer_state_dict"])
        for cur, saved in zip(self.fp32_groups, sd["fp32_groups"]):
            for c, s in zip(cur, saved):
                c.data.copy_(s.data)
