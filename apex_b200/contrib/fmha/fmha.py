"""Packed-QKV variable-length fused multi-head attention. Reference: apex/contrib/fmha/fmha.py:33-117 over ``fmhalib`` (sm_80
mma.sync kernels, fp16, head_dim 64, seq <= 512; deprecated upstream, removal July 2026).

Same interface — ``FMHA(config)(qkv [total, 3*hidden], cu_seqlens, max_s, is_training)`` — without the shape limits: any head
dim / sequence length / fp16-bf16-fp32. The attention core is online-softmax flash attention through
``torch.nn.functional.scaled_dot_product_attention`` on a padded view of the packed batch (library kernel, like cuBLAS for plain
GEMMs); a hand-written tcgen05/TMEM flash kernel is future work and is NOT claimed here."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _use_kernel(qkv: torch.Tensor, d: int, dropout: float) -> bool:
    """Opt-in (APEX_B200_FMHA_KERNEL=1) route through the experimental tcgen05 kernels: fp16 / bf16, head dim 64 or 128, no dropout."""
    from ...utils import config

    if not (config.fmha_kernel() and qkv.is_cuda and qkv.dtype in (torch.float16, torch.bfloat16) and d in (64, 128) and dropout == 0.0):
        return False
    from . import experimental as X

    return X.available()


def fmha_varlen(qkv: torch.Tensor, cu_seqlens: torch.Tensor, max_s: int, p_dropout: float = 0.0, is_training: bool = True,
                causal: bool = False) -> torch.Tensor:
    """qkv [total, 3, h, d] packed over sequences delimited by cu_seqlens [b+1] -> context [total, h, d]."""
    total, three, h, d = qkv.shape
    if _use_kernel(qkv, d, p_dropout if is_training else 0.0):
        from . import experimental as X

        cu = cu_seqlens if cu_seqlens.dtype == torch.int32 else cu_seqlens.to(torch.int32)
        ms = int(max_s) if max_s else int((cu[1:] - cu[:-1]).max())
        return X.FmhaFunc.apply(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu, cu, ms, ms, None, causal, None)
    b = cu_seqlens.numel() - 1
    lens = (cu_seqlens[1:] - cu_seqlens[:-1]).long()
    max_s = int(max_s) if max_s else int(lens.max())
    pos = torch.arange(max_s, device=qkv.device).unsqueeze(0)          # [1, S]
    valid = pos < lens.unsqueeze(1)                                    # [b, S]
    idx = (cu_seqlens[:-1].long().unsqueeze(1) + pos).clamp(max=total - 1)  # [b, S] gather indices into the packed dim
    padded = qkv[idx]                                                  # [b, S, 3, h, d]
    q, k, v = (padded[:, :, i].transpose(1, 2) for i in range(3))      # [b, h, S, d]
    mask = valid[:, None, None, :]                                     # keys beyond the length are masked
    if causal:
        mask = mask & torch.ones(max_s, max_s, dtype=torch.bool, device=qkv.device).tril()[None, None]
    out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=p_dropout if is_training else 0.0)
    out = out.transpose(1, 2)                                          # [b, S, h, d]
    return out[valid]                                                  # back to packed [total, h, d]


class FMHAFun(torch.autograd.Function):
    """Kept for API parity; autograd flows through :func:`fmha_varlen`."""

    @staticmethod
    def apply(qkv, cu_seqlens, p_dropout, max_s, is_training, zero_tensors=False):  # noqa: D102
        return fmha_varlen(qkv, cu_seqlens, max_s, p_dropout, is_training)


class FMHA(torch.nn.Module):
    def __init__(self, config):
        super().__init__()
        self.p_dropout = config.attention_probs_dropout_prob
        self.h = config.num_attention_heads
        self.hidden_size = config.hidden_size
        self.d = self.hidden_size // self.h
        assert self.d * self.h == self.hidden_size, "Invalid hidden size/num_heads"

    def forward(self, qkv, cu_seqlens, max_s, is_training=True, zero_tensors=False):
        ctx = fmha_varlen(qkv.view(-1, 3, self.h, self.d), cu_seqlens, max_s, self.p_dropout, is_training)
        return ctx.reshape(-1, self.hidden_size)
