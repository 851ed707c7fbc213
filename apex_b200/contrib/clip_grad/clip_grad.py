"""``clip_grad_norm_`` drop-in (reference apex/contrib/clip_grad/clip_grad.py:17-132): the 2-norm of CUDA gradients comes from the
multi-tensor L2-norm kernel (one launch per dtype, any of fp32/fp16/bf16) and the clip coefficient stays ON THE DEVICE — the scale
kernel reads it through a pointer, so there is no host synchronisation (the reference converts the coefficient tensor to a python
float for its kernel argument)."""
from __future__ import annotations

from typing import Iterable, Union

import torch

from ... import _lib
from ...ops import amp_C

_tensor_or_tensors = Union[torch.Tensor, Iterable[torch.Tensor]]


def clip_grad_norm_(parameters: _tensor_or_tensors, max_norm: float, norm_type: float = 2.0, error_if_nonfinite: bool = False) -> torch.Tensor:
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    parameters = [p for p in parameters if p.grad is not None]
    max_norm, norm_type = float(max_norm), float(norm_type)
    if len(parameters) == 0:
        return torch.tensor(0.0)
    if not (norm_type == 2.0 and any(p.is_cuda for p in parameters) and _lib.available()):
        return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=norm_type, error_if_nonfinite=error_if_nonfinite)
    device = next(p.device for p in parameters if p.is_cuda)
    buckets: dict = {}
    misc = []
    for p in parameters:
        g = p.grad.detach()
        if g.device == device and g.dtype in (torch.float32, torch.float16, torch.bfloat16) and g.is_contiguous():
            buckets.setdefault(g.dtype, []).append(g)
        else:
            misc.append(g)
    noop = torch.zeros(1, dtype=torch.int32, device=device)
    norms = [amp_C.multi_tensor_l2norm(65536, noop, [gs], False)[0] for gs in buckets.values()]
    norms += [torch.linalg.norm(g).unsqueeze(0).to(device) for g in misc]
    total_norm = torch.linalg.norm(torch.cat(norms))
    if error_if_nonfinite and torch.logical_or(total_norm.isnan(), total_norm.isinf()):
        raise RuntimeError(f"The total norm of order {norm_type} for gradients from `parameters` is non-finite, so it cannot be clipped. "
                           "To disable this error and scale the gradients by the non-finite norm anyway, set `error_if_nonfinite=False`")
    clip_coef_clamped = torch.clamp(max_norm / (total_norm + 1e-6), max=1.0).reshape(1)
    for gs in buckets.values():
        amp_C.multi_tensor_scale(65536, noop, [gs, gs], clip_coef_clamped)
    for g in misc:
        g.mul_(clip_coef_clamped.to(g.device))
    return total_norm
