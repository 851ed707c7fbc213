"""Plain-PyTorch fp32 implementations of every multi-tensor op.

Two jobs: (1) the CPU / no-extension code path (BASELINE.json config #1 "CPU python-only fallback"), and (2) the numerics
oracle the GPU tests compare the sm_100a kernels against. Semantics follow the reference functors cited in ops/amp_C.py.
"""
from __future__ import annotations

import math

import torch


def _finite_flag(noop_flag, *tensors):
    if noop_flag is None:
        return
    for t in tensors:
        if not torch.isfinite(t).all():
            noop_flag.fill_(1)
            return


def multi_tensor_scale(noop_flag, lists, scale):
    for i, o in zip(*lists):
        _finite_flag(noop_flag, i)
        o.copy_((i.float() * scale).to(o.dtype))


def multi_tensor_axpby(noop_flag, lists, a, b, arg_to_check):
    for x, y, o in zip(*lists):
        if arg_to_check == -1:
            _finite_flag(noop_flag, x, y)
        elif arg_to_check == 0:
            _finite_flag(noop_flag, x)
        else:
            _finite_flag(noop_flag, y)
        o.copy_((a * x.float() + b * y.float()).to(o.dtype))


def multi_tensor_l2norm(noop_flag, lists, per_tensor=False, inv_scale=None):
    xs = lists[0]
    dev = xs[0].device if xs else (noop_flag.device if noop_flag is not None else "cpu")
    s = 1.0 if inv_scale is None else inv_scale.float().reshape(())
    pts = [((x.float() * s) ** 2).sum().sqrt() for x in xs]
    pt = torch.stack(pts) if pts else torch.zeros(0, device=dev)
    tot = (pt ** 2).sum().sqrt().reshape(1) if pts else torch.zeros(1, device=dev)
    if noop_flag is not None and not torch.isfinite(tot).all():
        noop_flag.fill_(1)
    return tot, (pt if per_tensor else torch.empty(0, device=dev))


def multi_tensor_l2norm_scale(noop_flag, lists, scale, per_tensor=False):
    multi_tensor_scale(noop_flag, lists, scale)
    return multi_tensor_l2norm(None, [lists[1]], per_tensor)


def multi_tensor_norm_out(lists, out, alpha, beta, norm_type):
    for k, x in enumerate(lists[0]):
        if norm_type == 0:
            n = x.float().abs().max()
            out[k] = alpha * out[k] + beta * n
        else:
            n = (x.float() ** 2).sum().sqrt()
            out[k] = torch.sqrt(alpha * out[k] ** 2 + beta * n ** 2)


def _adam_math(g, p, m, v, lr, beta1, beta2, eps, bc1, bc2, mode, wd):
    if mode == 0:
        g = g + wd * p
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = (v / bc2).sqrt() + eps
    upd = (m / bc1) / denom
    if mode != 0:
        upd = upd + wd * p
    return p - lr * upd


def multi_tensor_adam(lists, lr, beta1, beta2, eps, step, mode, bias_correction, wd):
    bc1 = 1 - beta1 ** step if bias_correction else 1.0
    bc2 = 1 - beta2 ** step if bias_correction else 1.0
    for g, p, m, v in zip(*lists[:4]):
        mf, vf = m.float(), v.float()
        newp = _adam_math(g.float(), p.float(), mf, vf, lr, beta1, beta2, eps, bc1, bc2, mode, wd)
        m.copy_(mf)
        v.copy_(vf)
        p.copy_(newp.to(p.dtype))


def multi_tensor_adam_capturable(noop_flag, lists, lr, beta1, beta2, eps, step, mode, bias_correction, wd, inv_scale):
    if noop_flag is not None and int(noop_flag.item()) != 0:
        return
    s = float(step.item())
    bc1 = 1 - beta1 ** s if bias_correction else 1.0
    bc2 = 1 - beta2 ** s if bias_correction else 1.0
    inv = 1.0 if inv_scale is None else float(inv_scale.item())
    lrv = float(lr.item())
    master = lists[4] if len(lists) == 5 else None
    for k, (g, p, m, v) in enumerate(zip(*lists[:4])):
        gf = g.float() * inv
        g.copy_(gf.to(g.dtype))
        src = master[k] if master is not None else p
        mf, vf = m.float(), v.float()
        newp = _adam_math(gf, src.float(), mf, vf, lrv, beta1, beta2, eps, bc1, bc2, mode, wd)
        m.copy_(mf)
        v.copy_(vf)
        src.copy_(newp.to(src.dtype))
        if master is not None:
            p.copy_(newp.to(p.dtype))


def multi_tensor_adagrad(lists, lr, eps, mode, wd):
    for g, p, h in zip(*lists):
        gf, pf, hf = g.float(), p.float(), h.float()
        if mode == 0:
            gf = gf + wd * pf
            hf = hf + gf * gf
            pf = pf - lr * (gf / (hf.sqrt() + eps))
        else:
            hf = hf + gf * gf
            pf = pf - lr * (gf / (hf.sqrt() + eps) + wd * pf)
        h.copy_(hf.to(h.dtype))
        p.copy_(pf.to(p.dtype))


def multi_tensor_sgd(noop_flag, lists, wd, momentum, dampening, lr, nesterov, first_run, wd_after_momentum, scale):
    if noop_flag is not None and int(noop_flag.item()) != 0:
        return
    model = lists[3] if len(lists) == 4 else None
    for k, (g, p, mom) in enumerate(zip(*lists[:3])):
        gf, pf, mf = g.float() * scale, p.float(), mom.float()
        if wd != 0 and not wd_after_momentum:
            gf = gf + wd * pf
        if momentum != 0:
            mf = gf.clone() if first_run else mf * momentum + (1 - dampening) * gf
            gf = gf + momentum * mf if nesterov else mf
        if wd != 0 and wd_after_momentum:
            gf = gf + wd * pf
        pf = pf - lr * gf
        p.copy_(pf.to(p.dtype))
        if momentum != 0:
            mom.copy_(mf.to(mom.dtype))
        if model is not None:
            model[k].copy_(pf.to(model[k].dtype))


def multi_tensor_novograd(lists, grad_norms, lr, beta1, beta2, eps, step, bias_correction, wd, grad_averaging, moment_mode, norm_type):
    bc1 = 1 - beta1 ** step if bias_correction else 1.0
    bc2 = math.sqrt(1 - beta2 ** step) if bias_correction else 1.0
    beta3 = 1 - beta1 if grad_averaging else 1.0
    multi_tensor_norm_out([lists[0]], grad_norms, beta2, 1 - beta2, norm_type)
    for k, (g, p, m) in enumerate(zip(*lists)):
        gf, pf, mf = g.float(), p.float(), m.float()
        denom = grad_norms[k].float() / bc2 + eps
        if moment_mode == 0:
            gf = gf / denom + wd * pf
            mf = beta1 * mf + beta3 * gf
            pf = pf - lr * (mf / bc1)
        else:
            mf = beta1 * mf + beta3 * gf
            pf = pf - lr * ((mf / bc1) / denom + wd * pf)
        m.copy_(mf.to(m.dtype))
        p.copy_(pf.to(p.dtype))


def _lamb_core(lists, lr, beta1, beta2, beta3, eps, bc1, bc2, wd, mode, clip, use_nvlamb, inv_scale=1.0):
    model = lists[4] if len(lists) == 5 else None
    for k, (g, p, m, v) in enumerate(zip(*lists[:4])):
        gf, pf, mf, vf = g.float() * inv_scale / clip, p.float(), m.float(), v.float()
        if mode == 0:
            gf = gf + wd * pf
        mf = mf * beta1 + beta3 * gf
        vf = vf * beta2 + (1 - beta2) * gf * gf
        upd = (mf / bc1) / ((vf / bc2).sqrt() + eps)
        if mode != 0:
            upd = upd + wd * pf
        ratio = lr
        if use_nvlamb or wd != 0:
            pn, un = pf.norm(), upd.norm()
            if pn != 0 and un != 0:
                ratio = lr * float(pn / un)
        pf = pf - ratio * upd
        g.copy_(upd.to(g.dtype))
        m.copy_(mf.to(m.dtype))
        v.copy_(vf.to(v.dtype))
        p.copy_(pf.to(p.dtype))
        if model is not None:
            model[k].copy_(pf.to(model[k].dtype))


def multi_tensor_lamb(lists, lr, beta1, beta2, eps, step, bias_correction, wd, grad_averaging, mode, global_grad_norm, max_grad_norm,
                      use_nvlamb):
    bc1 = 1 - beta1 ** step if bias_correction else 1.0
    bc2 = 1 - beta2 ** step if bias_correction else 1.0
    beta3 = 1 - beta1 if grad_averaging else 1.0
    ggn = float(global_grad_norm.item()) if global_grad_norm is not None else 0.0
    clip = ggn / max_grad_norm if (max_grad_norm > 0 and ggn > max_grad_norm) else 1.0
    _lamb_core(lists, lr, beta1, beta2, beta3, eps, bc1, bc2, wd, mode, clip, use_nvlamb)


def multi_tensor_lamb_mp(noop_flag, lists, lr, beta1, beta2, eps, step, bias_correction, wd, grad_averaging, mode, global_grad_norm,
                         max_grad_norm, use_nvlamb, found_inf, inv_scale):
    if noop_flag is not None and int(noop_flag.item()) != 0:
        return
    s = float(step.item())
    bc1 = 1 - beta1 ** s if bias_correction else 1.0
    bc2 = 1 - beta2 ** s if bias_correction else 1.0
    beta3 = 1 - beta1 if grad_averaging else 1.0
    ggn, mx = float(global_grad_norm.item()), float(max_grad_norm.item())
    clip = ggn / mx if (mx > 0 and ggn > mx) else 1.0
    _lamb_core(lists, float(lr.item()), beta1, beta2, beta3, eps, bc1, bc2, wd, mode, clip, use_nvlamb, float(inv_scale.item()))


def multi_tensor_lamb_stage1(lists, per_tensor_decay, step, beta1, beta2, eps, global_grad_norm, max_global_grad_norm):
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    ggn = float(global_grad_norm.item())
    clip = ggn / max_global_grad_norm if ggn > max_global_grad_norm else 1.0
    for k, (g, p, m, v, u) in enumerate(zip(*lists)):
        gf, pf, mf, vf = g.float() / clip, p.float(), m.float(), v.float()
        mf = mf * beta1 + (1 - beta1) * gf
        vf = vf * beta2 + (1 - beta2) * gf * gf
        upd = (mf / bc1) / ((vf / bc2).sqrt() + eps) + float(per_tensor_decay[k]) * pf
        m.copy_(mf.to(m.dtype))
        v.copy_(vf.to(v.dtype))
        u.copy_(upd.to(u.dtype))


def multi_tensor_lamb_stage2(lists, pnorm, unorm, lr, wd, use_nvlamb):
    for k, (p, u) in enumerate(zip(*lists)):
        ratio = lr
        if use_nvlamb or wd != 0:
            pn, un = float(pnorm[k]), float(unorm[k])
            if pn != 0 and un != 0:
                ratio = lr * pn / un
        p.copy_((p.float() - ratio * u.float()).to(p.dtype))


def update_scale_hysteresis(scale, growth_tracker, hysteresis_tracker, found_inf, growth_factor, backoff_factor, growth_interval,
                            hysteresis):
    if float(found_inf.item()) > 0:
        hysteresis_tracker -= 1
        if int(hysteresis_tracker.item()) > 0:
            growth_tracker.zero_()
            return scale
        scale.mul_(backoff_factor)
        growth_tracker.zero_()
        return scale
    succ = int(growth_tracker.item()) + 1
    if succ == growth_interval:
        ns = scale * growth_factor
        if torch.isfinite(ns).all():
            scale.copy_(ns)
        growth_tracker.zero_()
    else:
        growth_tracker.fill_(succ)
    hysteresis_tracker.fill_(hysteresis)
    return scale


def dist_adam(p_in, m, v, g, p_out, grad_scale, lr, beta1, beta2, eps, step, mode, bias_correction, wd):
    """Sharded-optimizer Adam step on flat shards (lerp-form moments, reference distopt kernel)."""
    bc1 = 1 - beta1 ** step if bias_correction else 1.0
    bc2 = 1 - beta2 ** step if bias_correction else 1.0
    pf, mf, vf = p_in.float(), m.float(), v.float()
    sg = g.float() * float(grad_scale)
    if mode == 0:
        sg = sg + wd * pf
    mf = torch.lerp(sg, mf, beta1)
    vf = torch.lerp(sg * sg, vf, beta2)
    upd = (mf / bc1) / ((vf / bc2).sqrt() + eps)
    if mode != 0:
        upd = upd + wd * pf
    pf = pf - lr * upd
    m.copy_(mf.to(m.dtype))
    v.copy_(vf.to(v.dtype))
    p_in.copy_(pf.to(p_in.dtype))
    if p_out is not None and p_out.data_ptr() != p_in.data_ptr():
        p_out.copy_(pf.to(p_out.dtype))
