"""Fused sigmoid focal loss (detection). Reference: apex/contrib/focal_loss/focal_loss.py:6-69 over focal_loss_cuda
(forward returns the loss and stashes the un-normalised gradient; backward rescales it in place)."""
from __future__ import annotations

import torch

from ... import _lib

_lib.declare("ab_focal_loss_fwd", "p p p p p p l i i f f f i p")
_lib.declare("ab_focal_loss_bwd", "p p p l i p")


def _ref(cls_output, targets, num_positives_sum, num_real_classes, alpha, gamma, smoothing):
    x = cls_output.float()
    n_cls = x.shape[-1]
    x2 = x.reshape(-1, n_cls)
    t = targets.reshape(-1)
    onehot = torch.zeros_like(x2)
    pos = t >= 0
    onehot[pos, t[pos]] = 1.0
    y = onehot * (1 - smoothing) + smoothing * 0.5 if smoothing > 0 else onehot
    p = torch.sigmoid(x2)
    ce = torch.nn.functional.binary_cross_entropy_with_logits(x2, y, reduction="none")
    pt = torch.where(onehot > 0, 1 - p, p)
    at = torch.where(onehot > 0, torch.full_like(p, alpha), torch.full_like(p, 1 - alpha))
    loss = at * pt.pow(gamma) * ce
    valid = (t != -2).unsqueeze(1) & (torch.arange(n_cls, device=x.device) < num_real_classes).unsqueeze(0)
    return (loss * valid).sum() / num_positives_sum.float().reshape(())


class FocalLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_output, cls_targets_at_level, num_positives_sum, num_real_classes, alpha, gamma, label_smoothing=0.0):
        x = cls_output.contiguous()
        n_cls = x.shape[-1]
        n_ex = x.numel() // n_cls
        tgt = cls_targets_at_level.contiguous().view(-1).to(torch.int64)
        npos = num_positives_sum.float().reshape(1)
        pgrad = torch.empty_like(x)
        part = torch.empty(148 * 8, dtype=torch.float32, device=x.device)
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        _lib.fn("ab_focal_loss_fwd")(x.data_ptr(), tgt.data_ptr(), npos.data_ptr(), pgrad.data_ptr(), part.data_ptr(), loss.data_ptr(), n_ex,
                                     n_cls, int(num_real_classes), float(alpha), float(gamma), float(label_smoothing), _lib.dt(x),
                                     _lib.stream_ptr(x.device))
        ctx.save_for_backward(pgrad, npos)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        pgrad, npos = ctx.saved_tensors
        g = grad_loss.float().reshape(1).contiguous()
        _lib.fn("ab_focal_loss_bwd")(pgrad.data_ptr(), g.data_ptr(), npos.data_ptr(), pgrad.numel(), _lib.dt(pgrad), _lib.stream_ptr(pgrad.device))
        return pgrad, None, None, None, None, None, None


def focal_loss(cls_output, cls_targets_at_level, num_positive_sum, num_real_classes, alpha, gamma, label_smoothing=0.0):
    """cls_output [..., num_classes] logits; targets [...] class id, -1 background, -2 ignored; classes >= num_real_classes are padding."""
    if cls_output.is_cuda and _lib.available() and cls_output.dtype in (torch.float32, torch.float16, torch.bfloat16):
        return FocalLoss.apply(cls_output, cls_targets_at_level, num_positive_sum, num_real_classes, alpha, gamma, label_smoothing)
    return _ref(cls_output, cls_targets_at_level, num_positive_sum, num_real_classes, alpha, gamma, label_smoothing)
