"""Packed-QKV variable-length fused multi-head attention. Reference: apex/contrib/fmha/fmha.py:33-117 over ``fmhalib`` (sm_80
mma.sync kernels, fp16, head_dim 64, seq <= 512; deprecated upstream, removal July 2026).

Same interface — ``FMHA(config)(qkv [total, 3*hidden], cu_seqlens, max_s, is_training)`` — without the sequence-length limit.
fp16 / bf16 inputs with head dim 64 or 128 run on the hand-written tcgen05 / TMEM / TMA kernels (csrc/fmha_fwd_sm100.cu,
csrc/fmha_bwd_sm100.cu through :mod:`.kernels`): varlen batches straight from the packed layout (3-D TMA maps over
[rows, heads, d], no padding pass), causal masking, Philox dropout regenerated in the backward, deterministic (atomic-free)
gradients. Other dtypes / head dims (fp32, d = 96, ...) compose the generic path: padded view + the repo's scaled-masked-softmax
kernel between two batched GEMMs."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import kernels as K


def _generic_varlen(qkv, cu_seqlens, max_s, p_dropout, causal):
    """Any dtype / head dim: pad to [b, h, S, d], scores by batched GEMM, the fused scaled-masked-softmax kernel, context by batched GEMM."""
    from ...transformer.functional import scaled_masked_softmax

    total, _, h, d = qkv.shape
    b = cu_seqlens.numel() - 1
    lens = (cu_seqlens[1:] - cu_seqlens[:-1]).long()
    max_s = int(max_s) if max_s else int(lens.max())
    pos = torch.arange(max_s, device=qkv.device).unsqueeze(0)          # [1, S]
    valid = pos < lens.unsqueeze(1)                                    # [b, S]
    idx = (cu_seqlens[:-1].long().unsqueeze(1) + pos).clamp(max=total - 1)  # [b, S] gather indices into the packed dim
    padded = qkv[idx]                                                  # [b, S, 3, h, d]
    q, k, v = (padded[:, :, i].transpose(1, 2) for i in range(3))      # [b, h, S, d]
    masked = ~valid[:, None, None, :].expand(b, 1, max_s, max_s)       # True = masked (keys beyond the length)
    if causal:
        masked = masked | torch.ones(max_s, max_s, dtype=torch.bool, device=qkv.device).triu(1)[None, None]
    scores = torch.matmul(q, k.transpose(-1, -2))
    p = scaled_masked_softmax(scores, masked.to(torch.uint8), d ** -0.5)
    p = F.dropout(p, p_dropout, p_dropout > 0.0)
    out = torch.matmul(p.to(v.dtype), v).transpose(1, 2)               # [b, S, h, d]
    return out[valid]                                                  # back to packed [total, h, d]


def fmha_varlen(qkv: torch.Tensor, cu_seqlens: torch.Tensor, max_s: int, p_dropout: float = 0.0, is_training: bool = True,
                causal: bool = False) -> torch.Tensor:
    """qkv [total, 3, h, d] packed over sequences delimited by cu_seqlens [b+1] -> context [total, h, d]."""
    total, three, h, d = qkv.shape
    p = float(p_dropout) if is_training else 0.0
    if K.supported(qkv, d):
        cu = cu_seqlens if cu_seqlens.dtype == torch.int32 else cu_seqlens.to(torch.int32)
        ms = int(max_s) if max_s else int((cu[1:] - cu[:-1]).max())
        return K.FmhaFunc.apply(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu, cu, ms, ms, None, causal, None, None, p)
    if qkv.is_cuda and qkv.dtype in (torch.float16, torch.bfloat16) and d in (64, 128):
        raise RuntimeError("apex_b200.contrib.fmha: the native library is required on a CUDA device (python -m apex_b200._build)")
    return _generic_varlen(qkv, cu_seqlens, max_s, p, causal)


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
