"""Pure-PyTorch oracles with the reference's helper signatures (apex/contrib/transducer/_transducer_ref.py:4-118): used by tests that
compare the fused kernels against a lattice computed the slow, obvious way."""
from __future__ import annotations

import torch

from .transducer import _TorchTransducerJoint, _TorchTransducerLoss


def transducer_loss_reference(x, label, f_len, y_len, blank_idx, loss_grad):
    """x [B, T, U, V] logits (requires_grad) -> (alpha, beta, x.grad, loss): forward / backward variables of the RNN-T lattice in log
    space (cells outside an utterance's f_len x (y_len + 1) rectangle are 0), the gradient for upstream ``loss_grad`` and -log p(y|x)."""
    B, T, U, _ = x.shape
    lp = torch.log_softmax(x.detach().float(), dim=-1)
    alpha, beta = torch.zeros(B, T, U), torch.zeros(B, T, U)
    for b in range(B):
        Tb, Ub = int(f_len[b]), int(y_len[b]) + 1
        blank = lp[b, :, :, blank_idx]
        emit = lp[b, :, :U - 1, :].gather(2, label[b].long().clamp(min=0).view(1, U - 1, 1).expand(T, U - 1, 1)).squeeze(2)
        for t in range(Tb):
            for u in range(Ub):
                if t == 0 and u == 0:
                    continue
                terms = []
                if t > 0:
                    terms.append(alpha[b, t - 1, u] + blank[t - 1, u])
                if u > 0:
                    terms.append(alpha[b, t, u - 1] + emit[t, u - 1])
                alpha[b, t, u] = torch.logsumexp(torch.stack(terms), 0)
        for t in range(Tb - 1, -1, -1):
            for u in range(Ub - 1, -1, -1):
                if t == Tb - 1 and u == Ub - 1:
                    beta[b, t, u] = blank[t, u]
                    continue
                terms = []
                if t < Tb - 1:
                    terms.append(beta[b, t + 1, u] + blank[t, u])
                if u < Ub - 1:
                    terms.append(beta[b, t, u + 1] + emit[t, u])
                beta[b, t, u] = torch.logsumexp(torch.stack(terms), 0)
    loss = _TorchTransducerLoss()(x, label, f_len, y_len, blank_idx)
    loss.backward(loss_grad.to(loss.dtype))
    return alpha.to(x.device), beta.to(x.device), x.grad, loss.detach().to(x.dtype)


def transducer_joint_reference(f, g, h_grad, f_len, g_len, pack_output, relu, dropout, dropout_prob=0, mask=None):
    """f [B, T, H], g [B, U, H] (both requires_grad) -> (h, f.grad, g.grad); with ``dropout`` the caller supplies the keep-mask."""
    if dropout and mask is None:
        raise NotImplementedError("mask needs to supplied to test dropout.")
    h = f.unsqueeze(2) + g.unsqueeze(1)
    if relu:
        h = torch.relu(h)
    if dropout:
        h = h * mask.to(h.dtype) / (1.0 - dropout_prob)
    B, T, U, _ = h.shape
    valid = (torch.arange(T, device=f.device).view(1, T, 1) < f_len.view(B, 1, 1)) & (torch.arange(U, device=f.device).view(1, 1, U) < g_len.view(B, 1, 1))
    out = h[valid] if pack_output else h * valid.unsqueeze(-1).to(h.dtype)
    out.backward(h_grad)
    return out.detach(), f.grad, g.grad


__all__ = ["transducer_loss_reference", "transducer_joint_reference", "_TorchTransducerJoint", "_TorchTransducerLoss"]
