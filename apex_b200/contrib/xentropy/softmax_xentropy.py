"""Fused softmax cross-entropy with label smoothing (csrc/xentropy.cu). Reference: apex/contrib/xentropy/softmax_xentropy.py:6-33.
Only ``max + log-sum-exp`` (one float per row) is saved; the backward recomputes the softmax and may write the gradient in
place over the logits (``inplace_backward=True``, the reference behaviour)."""
from __future__ import annotations

import torch

from ... import _lib

_lib.declare("ab_xentropy_fwd", "p p p p i i f l i p")
_lib.declare("ab_xentropy_bwd", "p p p p p i i f l i p")


class SoftmaxCrossEntropyLoss(torch.autograd.Function):
    inplace_backward = False

    @staticmethod
    def forward(ctx, logits, labels, smoothing=0.0, padding_idx=0, half_to_float=False):
        x = logits.contiguous()
        rows, C = x.numel() // x.shape[-1], x.shape[-1]
        lab = labels.contiguous().view(-1).to(torch.int64)
        if x.is_cuda and _lib.available() and x.dtype in (torch.float16, torch.bfloat16, torch.float32):
            losses = torch.empty(rows, dtype=torch.float32, device=x.device)
            mlse = torch.empty(rows, dtype=torch.float32, device=x.device)
            _lib.fn("ab_xentropy_fwd")(x.data_ptr(), lab.data_ptr(), losses.data_ptr(), mlse.data_ptr(), rows, C, float(smoothing),
                                       int(padding_idx), _lib.dt(x), _lib.stream_ptr(x.device))
        else:
            xf = x.float().view(rows, C)
            mlse = torch.logsumexp(xf, dim=-1)
            xl = xf.gather(1, lab.clamp(0, C - 1).view(-1, 1)).squeeze(1)
            losses = (mlse - xf.mean(-1)) * smoothing - (xl - mlse) * (1.0 - smoothing)
            losses = losses.masked_fill(lab == padding_idx, 0.0)
        ctx.save_for_backward(x, mlse, lab)
        ctx.smoothing, ctx.padding_idx, ctx.shape = smoothing, padding_idx, labels.shape
        out = losses.view(labels.shape)
        return out if (half_to_float or x.dtype == torch.float32) else out.to(x.dtype)

    @staticmethod
    def backward(ctx, grad_loss):
        x, mlse, lab = ctx.saved_tensors
        rows, C = x.numel() // x.shape[-1], x.shape[-1]
        g = grad_loss.contiguous().view(-1).float()
        if x.is_cuda and _lib.available() and x.dtype in (torch.float16, torch.bfloat16, torch.float32):
            gl = x if SoftmaxCrossEntropyLoss.inplace_backward else torch.empty_like(x)
            _lib.fn("ab_xentropy_bwd")(g.data_ptr(), x.data_ptr(), mlse.data_ptr(), lab.data_ptr(), gl.data_ptr(), rows, C,
                                       float(ctx.smoothing), int(ctx.padding_idx), _lib.dt(x), _lib.stream_ptr(x.device))
        else:
            xf = x.float().view(rows, C)
            p = torch.exp(xf - mlse.view(-1, 1))
            oh = torch.zeros_like(p).scatter_(1, lab.clamp(0, C - 1).view(-1, 1), 1.0)
            gl = (g.masked_fill(lab == ctx.padding_idx, 0.0).view(-1, 1) * (p - (1 - ctx.smoothing) * oh - ctx.smoothing / C)).to(x.dtype).view(x.shape)
        return gl, None, None, None, None
