"""`amp_C`-compatible multi-tensor ops (reference: csrc/amp_C_frontend.cpp:83-123) on the B200 device-table engine.

Every function takes ``(chunk_size, noop_flag, tensor_lists, *args)`` like the reference pybind functions, so
``multi_tensor_applier(amp_C.multi_tensor_adam, noop, lists, ...)`` keeps working. ``tensor_lists`` may also be a prebuilt
:class:`TensorTable` (the cached fast path the optimizers use: no per-step list walking / validation / re-packing).

CUDA tensors run the hand-written sm_100a kernels in ``csrc/mt_*.cu``; CPU tensors run the PyTorch reference
implementation (``ops/reference.py``), which doubles as the numerics oracle in the tests.
"""
from __future__ import annotations

import torch

from .. import _lib
from . import reference as ref

_scratch: dict[tuple, torch.Tensor] = {}


def _partials(device, n: int, slot: int = 0) -> torch.Tensor:
    key = (device, slot)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < n:
        buf = _scratch[key] = torch.empty(max(n, 4096), dtype=torch.float32, device=device)
    return buf


class TensorTable:
    """Device-resident descriptor of ``depth`` parallel tensor lists (see csrc/binding.cpp)."""

    def __init__(self, tensor_lists, chunk_size: int = 65536):
        self._t = _lib.host_runtime().TensorTable()
        self._t.build([list(l) for l in tensor_lists], int(chunk_size))
        self.device = tensor_lists[0][0].device if len(tensor_lists[0]) else None

    def track_grads(self, grad_slot: int, param_slot: int) -> None:
        self._t.track_grads(grad_slot, param_slot)

    def track_universe(self, params, in_table) -> None:
        self._t.track_universe(list(params), [bool(b) for b in in_table])

    def refresh_grads(self) -> bool:
        return self._t.refresh_grads()

    def set_slot(self, d: int, tensors) -> None:
        self._t.set_slot(d, list(tensors))

    def slot(self, d: int):
        return self._t.slot(d)

    @property
    def n(self):
        return self._t.n

    @property
    def depth(self):
        return self._t.depth

    @property
    def total_chunks(self):
        return self._t.total_chunks

    @property
    def uploads(self):
        return self._t.uploads

    @property
    def dtypes(self):
        return self._t.dtypes

    @property
    def total_numel(self):
        return self._t.total_numel

    def head(self):
        t = self._t
        return (t.arena, t.n, t.depth, t.total_chunks, t.chunk)


def _is_cuda(lists) -> bool:
    if isinstance(lists, TensorTable):
        return True
    return len(lists) > 0 and len(lists[0]) > 0 and lists[0][0].is_cuda


def _table(lists, chunk_size) -> TensorTable:
    if isinstance(lists, TensorTable):
        return lists
    if not _lib.available():
        raise _lib.gpu_required_error("multi_tensor_apply")
    return TensorTable(lists, chunk_size if chunk_size and chunk_size > 0 else 65536)   # optimizers pass 0 together with prebuilt tables


def _empty(lists) -> bool:
    if isinstance(lists, TensorTable):
        return lists.n == 0
    return len(lists) == 0 or len(lists[0]) == 0


def _s(tb: TensorTable) -> int:
    return _lib.stream_ptr(tb.device)


# ---------------------------------------------------------------------------------------------------------------------
def multi_tensor_scale(chunk_size, noop_flag, tensor_lists, scale):
    """out = in * scale. ``scale`` may be a python number or a 1-element DEVICE tensor (no host sync, graph-capturable)."""
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_scale(noop_flag, tensor_lists, float(scale))
    tb = _table(tensor_lists, chunk_size)
    d = tb.dtypes
    if torch.is_tensor(scale):
        sp = scale.to(device=tb.device, dtype=torch.float32)
        _lib.fn("ab_mt_scale")(*tb.head(), d[0], d[1], 1.0, _lib.ptr(noop_flag), sp.data_ptr(), _s(tb))
    else:
        _lib.fn("ab_mt_scale")(*tb.head(), d[0], d[1], float(scale), _lib.ptr(noop_flag), None, _s(tb))


def multi_tensor_axpby(chunk_size, noop_flag, tensor_lists, a, b, arg_to_check):
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_axpby(noop_flag, tensor_lists, a, b, arg_to_check)
    tb = _table(tensor_lists, chunk_size)
    d = tb.dtypes
    _lib.fn("ab_mt_axpby")(*tb.head(), d[0], d[1], d[2], float(a), float(b), int(arg_to_check), _lib.ptr(noop_flag), _s(tb))


def _norm(chunk_size, noop_flag, tensor_lists, per_tensor, *, inv_scale=None, noop_in=None, is_max=False, out_per_tensor=None,
          blend=None, want_global=True):
    tb = _table(tensor_lists, chunk_size)
    dev = tb.device
    part = _partials(dev, tb.total_chunks)
    out = torch.empty(1, dtype=torch.float32, device=dev) if want_global else None
    pt = out_per_tensor
    if pt is None and per_tensor:
        pt = torch.empty(tb.n, dtype=torch.float32, device=dev)
    a, b = blend if blend is not None else (0.0, 0.0)
    _lib.fn("ab_mt_norm")(*tb.head(), tb.dtypes[0], part.data_ptr(), _lib.ptr(out), _lib.ptr(pt), _lib.ptr(inv_scale),
                          _lib.ptr(noop_in), _lib.ptr(noop_flag), int(is_max), int(blend is not None), float(a), float(b), _s(tb))
    return out, pt


def multi_tensor_l2norm(chunk_size, noop_flag, tensor_lists, per_tensor=False):
    """-> (global_norm[1], per_tensor_norms[n] or empty)."""
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_l2norm(noop_flag, tensor_lists, per_tensor)
    if _empty(tensor_lists):
        dev = noop_flag.device
        return torch.zeros(1, device=dev), torch.empty(0, device=dev)
    out, pt = _norm(chunk_size, noop_flag, tensor_lists, per_tensor)
    return out, (pt if pt is not None else torch.empty(0, dtype=torch.float32, device=out.device))


def multi_tensor_unscale_l2norm(chunk_size, noop_flag, tensor_lists, inv_scale, per_tensor=False):
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_l2norm(noop_flag, tensor_lists, per_tensor, inv_scale=inv_scale)
    if _empty(tensor_lists):
        dev = noop_flag.device
        return torch.zeros(1, device=dev), torch.empty(0, device=dev)
    out, pt = _norm(chunk_size, noop_flag, tensor_lists, per_tensor, inv_scale=inv_scale)
    return out, (pt if pt is not None else torch.empty(0, dtype=torch.float32, device=out.device))


def multi_tensor_l2norm_mp(chunk_size, noop_flag, tensor_lists, per_tensor=False):
    """l2norm that does nothing when ``noop_flag`` is already set (mixed-precision LAMB)."""
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_l2norm(noop_flag, tensor_lists, per_tensor)
    if _empty(tensor_lists):
        dev = noop_flag.device
        return torch.zeros(1, device=dev), torch.empty(0, device=dev)
    tb = _table(tensor_lists, chunk_size)
    part = _partials(tb.device, tb.total_chunks)
    out = torch.zeros(1, dtype=torch.float32, device=tb.device)
    pt = torch.zeros(tb.n, dtype=torch.float32, device=tb.device) if per_tensor else None
    _lib.fn("ab_mt_norm")(*tb.head(), tb.dtypes[0], part.data_ptr(), out.data_ptr(), _lib.ptr(pt), None, noop_flag.data_ptr(),
                          None, 0, 0, 0.0, 0.0, _s(tb))
    return out, (pt if pt is not None else torch.empty(0, dtype=torch.float32, device=out.device))


def multi_tensor_l2norm_scale(chunk_size, noop_flag, tensor_lists, scale, per_tensor=False):
    """[in, out]: out = in*scale, returns the L2 norm of ``out``."""
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_l2norm_scale(noop_flag, tensor_lists, scale, per_tensor)
    if _empty(tensor_lists):
        dev = noop_flag.device
        return torch.zeros(1, device=dev), torch.empty(0, device=dev)
    tb = _table(tensor_lists, chunk_size)
    part = _partials(tb.device, tb.total_chunks)
    out = torch.empty(1, dtype=torch.float32, device=tb.device)
    pt = torch.empty(tb.n, dtype=torch.float32, device=tb.device) if per_tensor else None
    d = tb.dtypes
    _lib.fn("ab_mt_l2norm_scale")(*tb.head(), d[0], d[1], float(scale), part.data_ptr(), out.data_ptr(), _lib.ptr(pt),
                                  _lib.ptr(noop_flag), _s(tb))
    return out, (pt if pt is not None else torch.empty(0, dtype=torch.float32, device=out.device))


def multi_tensor_norm_out(chunk_size, noop_flag, tensor_lists, out, alpha, beta, norm_type):
    """Per-tensor norm blended into the running ``out``: L2 (norm_type 2) sqrt(a*o^2+b*n^2); Linf (0) a*o+b*n."""
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_norm_out(tensor_lists, out, alpha, beta, norm_type)
    _norm(chunk_size, None, tensor_lists, True, is_max=(norm_type == 0), out_per_tensor=out, blend=(alpha, beta), want_global=False)


def multi_tensor_adam(chunk_size, noop_flag, tensor_lists, lr, beta1, beta2, eps, step, mode, bias_correction, weight_decay):
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_adam(tensor_lists, lr, beta1, beta2, eps, step, mode, bias_correction, weight_decay)
    tb = _table(tensor_lists, chunk_size)
    d = tb.dtypes
    _lib.fn("ab_mt_adam")(*tb.head(), d[0], d[1], float(lr), float(beta1), float(beta2), float(eps), int(step), int(mode),
                          int(bias_correction), float(weight_decay), 0, None, None, None, None, _s(tb))


def multi_tensor_adam_capturable(chunk_size, noop_flag, tensor_lists, lr, beta1, beta2, eps, step, mode, bias_correction,
                                 weight_decay, inv_scale):
    """lr / step / inv_scale are device tensors; skipped entirely when ``noop_flag`` is set. Grads are unscaled in place."""
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_adam_capturable(noop_flag, tensor_lists, lr, beta1, beta2, eps, step, mode, bias_correction,
                                                weight_decay, inv_scale)
    tb = _table(tensor_lists, chunk_size)
    d = tb.dtypes
    _lib.fn("ab_mt_adam")(*tb.head(), d[0], d[1], 0.0, float(beta1), float(beta2), float(eps), 0, int(mode), int(bias_correction),
                          float(weight_decay), 1, lr.data_ptr(), step.data_ptr(), _lib.ptr(inv_scale), _lib.ptr(noop_flag), _s(tb))


multi_tensor_adam_capturable_master = multi_tensor_adam_capturable  # 5 lists: [g, p, m, v, p_master]


def multi_tensor_adam_swa(chunk_size, tensor_lists, lr, beta1, beta2, eps, step, mode, torch_math, bias_correction, weight_decay, swa_a, swa_b,
                          clip_scale=None):
    """[g, p, m, v, swa, compute]: Adam on the fp32 parameters, swa = swa_a * swa + swa_b * p and the low-precision compute copy, one
    persistent launch (csrc/mt_optim.cu AdamSwaOp). ``clip_scale``: optional 1-element device tensor multiplied into the gradients."""
    if _empty(tensor_lists):
        return
    tb = _table(tensor_lists, chunk_size)
    d = tb.dtypes
    _lib.fn("ab_mt_adam_swa")(*tb.head(), d[0], d[5], float(lr), float(beta1), float(beta2), float(eps), int(step), int(mode), int(torch_math),
                              int(bias_correction), float(weight_decay), float(swa_a), float(swa_b), _lib.ptr(clip_scale), _s(tb))


def multi_tensor_adagrad(chunk_size, noop_flag, tensor_lists, lr, eps, mode, weight_decay):
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_adagrad(tensor_lists, lr, eps, mode, weight_decay)
    tb = _table(tensor_lists, chunk_size)
    _lib.fn("ab_mt_adagrad")(*tb.head(), tb.dtypes[0], float(lr), float(eps), int(mode), float(weight_decay), _s(tb))


def multi_tensor_sgd(chunk_size, noop_flag, tensor_lists, wd, momentum, dampening, lr, nesterov, first_run, wd_after_momentum,
                     scale=1.0):
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_sgd(noop_flag, tensor_lists, wd, momentum, dampening, lr, nesterov, first_run, wd_after_momentum, scale)
    tb = _table(tensor_lists, chunk_size)
    d = tb.dtypes
    dm = d[3] if tb.depth == 4 else 0
    _lib.fn("ab_mt_sgd")(*tb.head(), d[0], d[1], dm, float(wd), float(momentum), float(dampening), float(lr), int(nesterov),
                         int(first_run), int(wd_after_momentum), float(scale), _lib.ptr(noop_flag), _s(tb))


def multi_tensor_novograd(chunk_size, noop_flag, tensor_lists, grad_norms, lr, beta1, beta2, eps, step, bias_correction,
                          weight_decay, grad_averaging, moment_mode, norm_type):
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_novograd(tensor_lists, grad_norms, lr, beta1, beta2, eps, step, bias_correction, weight_decay,
                                         grad_averaging, moment_mode, norm_type)
    tb = _table(tensor_lists, chunk_size)
    # 1) blend this step's per-tensor grad norm into the running second moment, 2) update
    _norm(chunk_size, None, tb, True, is_max=(norm_type == 0), out_per_tensor=grad_norms, blend=(beta2, 1.0 - beta2), want_global=False)
    _lib.fn("ab_mt_novograd")(*tb.head(), tb.dtypes[0], float(lr), float(beta1), float(beta2), float(eps), int(step),
                              int(bias_correction), float(weight_decay), int(grad_averaging), int(moment_mode),
                              grad_norms.data_ptr(), _s(tb))


def _lamb(tb: TensorTable, *, beta1, beta2, beta3, step, bias_correction, eps, mode, decay, global_grad_norm, max_grad_norm,
          lr, use_nvlamb, device_scalars=False, step_ptr=None, inv_scale=None, noop=None, lr_ptr=None, max_norm_ptr=None,
          model_copy=False):
    dev = tb.device
    pp, pu = _partials(dev, tb.total_chunks, 1), _partials(dev, tb.total_chunks, 2)
    pn = torch.empty(tb.n, dtype=torch.float32, device=dev)
    un = torch.empty(tb.n, dtype=torch.float32, device=dev)
    d = tb.dtypes
    s = _s(tb)
    _lib.fn("ab_mt_lamb_stage1")(*tb.head(), d[0], d[1], 0, float(beta1), float(beta2), float(beta3), int(step), int(bias_correction),
                                 float(eps), int(mode), float(decay), None, _lib.ptr(global_grad_norm), float(max_grad_norm),
                                 _lib.ptr(max_norm_ptr), int(device_scalars), _lib.ptr(step_ptr), _lib.ptr(inv_scale), _lib.ptr(noop),
                                 pp.data_ptr(), pu.data_ptr(), pn.data_ptr(), un.data_ptr(), 0, s)
    _lib.fn("ab_mt_lamb_stage2")(*tb.head(), d[0], d[1], pn.data_ptr(), un.data_ptr(), float(lr), _lib.ptr(lr_ptr), float(decay), None,
                                 int(bool(use_nvlamb)), int(device_scalars), _lib.ptr(noop), int(model_copy), 0, s)


def multi_tensor_lamb(chunk_size, noop_flag, tensor_lists, lr, beta1, beta2, eps, step, bias_correction, weight_decay,
                      grad_averaging, mode, global_grad_norm, max_grad_norm, use_nvlamb=False):
    """[g, p, m, v]. Two passes: (update term + both per-tensor norms) then (trust-ratio apply). g is clobbered."""
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_lamb(tensor_lists, lr, beta1, beta2, eps, step, bias_correction, weight_decay, grad_averaging, mode,
                                     global_grad_norm, max_grad_norm, use_nvlamb)
    tb = _table(tensor_lists, chunk_size)
    beta3 = 1.0 - beta1 if grad_averaging else 1.0
    _lamb(tb, beta1=beta1, beta2=beta2, beta3=beta3, step=step, bias_correction=bias_correction, eps=eps, mode=mode,
          decay=weight_decay, global_grad_norm=global_grad_norm, max_grad_norm=max_grad_norm, lr=lr, use_nvlamb=use_nvlamb)


def multi_tensor_lamb_mp(chunk_size, noop_flag, tensor_lists, lr, beta1, beta2, eps, step, bias_correction, weight_decay,
                         grad_averaging, mode, global_grad_norm, max_grad_norm, use_nvlamb, found_inf, inv_scale):
    """Mixed-precision LAMB: lr/step/max_grad_norm/inv_scale are device tensors; 4 lists, or 5 with a low-precision model copy."""
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_lamb_mp(noop_flag, tensor_lists, lr, beta1, beta2, eps, step, bias_correction, weight_decay,
                                        grad_averaging, mode, global_grad_norm, max_grad_norm, use_nvlamb, found_inf, inv_scale)
    tb = _table(tensor_lists, chunk_size)
    beta3 = 1.0 - beta1 if grad_averaging else 1.0
    _lamb(tb, beta1=beta1, beta2=beta2, beta3=beta3, step=0, bias_correction=bias_correction, eps=eps, mode=mode,
          decay=weight_decay, global_grad_norm=global_grad_norm, max_grad_norm=1.0, lr=0.0, use_nvlamb=use_nvlamb,
          device_scalars=True, step_ptr=step, inv_scale=inv_scale, noop=noop_flag, lr_ptr=lr, max_norm_ptr=max_grad_norm,
          model_copy=(tb.depth == 5))


def multi_tensor_lamb_stage1_cuda(chunk_size, noop_flag, tensor_lists, per_tensor_decay, step, beta1, beta2, eps,
                                  global_grad_norm, max_global_grad_norm):
    """Legacy split LAMB, stage 1: [g, p, m, v, update] with a per-tensor decay tensor."""
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_lamb_stage1(tensor_lists, per_tensor_decay, step, beta1, beta2, eps, global_grad_norm, max_global_grad_norm)
    tb = _table(tensor_lists, chunk_size)
    dev = tb.device
    pp, pu = _partials(dev, tb.total_chunks, 1), _partials(dev, tb.total_chunks, 2)
    d = tb.dtypes
    _lib.fn("ab_mt_lamb_stage1")(*tb.head(), d[0], d[1], d[4], float(beta1), float(beta2), float(1.0 - beta1), int(step), 1, float(eps),
                                 1, 0.0, per_tensor_decay.data_ptr(), _lib.ptr(global_grad_norm), float(max_global_grad_norm), None, 0,
                                 None, None, None, pp.data_ptr(), pu.data_ptr(), None, None, 4, _s(tb))


def multi_tensor_lamb_stage2_cuda(chunk_size, noop_flag, tensor_lists, per_tensor_param_norm, per_tensor_update_norm, lr,
                                  weight_decay=1.0, use_nvlamb=True):
    """Legacy split LAMB, stage 2: [p, update]."""
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        return ref.multi_tensor_lamb_stage2(tensor_lists, per_tensor_param_norm, per_tensor_update_norm, lr, weight_decay, use_nvlamb)
    tb = _table(tensor_lists, chunk_size)
    d = tb.dtypes
    _lib.fn("ab_mt_lamb_stage2")(*tb.head(), d[1], d[0], per_tensor_param_norm.data_ptr(), per_tensor_update_norm.data_ptr(), float(lr),
                                 None, float(weight_decay), None, int(bool(use_nvlamb)), 0, None, 0, 1, _s(tb))


def multi_tensor_cast(chunk_size, noop_flag, tensor_lists, scale=1.0):
    """[in, out]: out = cast(in*scale); fp32/fp16/bf16/e5m2/e4m3 on either side (maybe_cast_mt equivalent)."""
    if _empty(tensor_lists):
        return
    if not _is_cuda(tensor_lists):
        for i, o in zip(*tensor_lists):
            o.copy_((i.float() * scale).to(o.dtype))
        return
    tb = _table(tensor_lists, chunk_size)
    d = tb.dtypes
    _lib.fn("ab_mt_cast")(*tb.head(), d[0], d[1], float(scale), _s(tb))


def update_scale_hysteresis(current_scale, growth_tracker, hysteresis_tracker, found_inf, growth_factor, backoff_factor,
                            growth_interval, hysteresis):
    if not current_scale.is_cuda:
        return ref.update_scale_hysteresis(current_scale, growth_tracker, hysteresis_tracker, found_inf, growth_factor, backoff_factor,
                                           growth_interval, hysteresis)
    if not _lib.available():
        raise _lib.gpu_required_error("update_scale_hysteresis")
    _lib.fn("ab_update_scale_hysteresis")(current_scale.data_ptr(), growth_tracker.data_ptr(), hysteresis_tracker.data_ptr(),
                                          found_inf.data_ptr(), float(growth_factor), float(backoff_factor), int(growth_interval),
                                          int(hysteresis), _lib.stream_ptr(current_scale.device))
    return current_scale
