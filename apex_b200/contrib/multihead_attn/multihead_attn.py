"""Self / encoder-decoder multi-head attention modules. Reference: apex/contrib/multihead_attn/{self,encdec}_multihead_attn.py and
the 16 ``fast_multihead_attn`` entry points (multihead_attn_frontend.cpp:573-605: cuBLAS strided-batched GEMMs / CUTLASS 1.x +
softmax+dropout kernels, optional pre-LayerNorm + residual add + dropout).

Same constructor / forward contract ([time, batch, channel] inputs, ``key_padding_mask`` or ``attn_mask``, additive or boolean masks,
``include_norm_add`` pre-LN residual variant, ``impl`` in {"fast", "default"}). On B200 both impls run: input/output projections on
the tcgen05 GEMM (apex_b200.ops.gemm through fused_dense_function), pre-LN on the fused LayerNorm kernel, and the attention core
(scores, key-padding / causal masking, softmax, dropout, context) in ONE tcgen05 / TMEM kernel per direction (contrib/fmha/kernels.py,
csrc/fmha_{fwd,bwd}_sm100.cu) that reads the packed projection output in place. fp32 inputs, head dims other than 64 / 128 and
arbitrary (non-causal) time masks compose the generic path: batched GEMM + the fused scaled-masked-softmax kernel + batched GEMM.

Parameter layout is the reference's, so its checkpoints load unchanged: the packed ``in_proj_weight`` ([3 * embed, embed]) is interleaved per
head — rows ordered [head][q | k | v][head_dim] — and ``in_proj_weight_kv`` likewise [head][k | v][head_dim]; ``separate_qkv_params``
holds three ordinary [embed, embed] matrices. :func:`packed_to_blocked` converts to the [q; k; v] block order of ``torch.nn.MultiheadAttention``."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Parameter

from ...fused_dense import fused_dense_function
from ...normalization import FusedLayerNorm
from ...normalization.fused_layer_norm import fused_layer_norm_affine
from ...transformer.functional import scaled_masked_softmax, scaled_softmax


def packed_to_blocked(t: torch.Tensor, heads: int, parts: int = 3) -> torch.Tensor:
    """Per-head interleaved projection weight / bias ([head][part][head_dim] rows) -> block order ([part][head][head_dim])."""
    rest = t.shape[1:]
    return t.reshape(heads, parts, t.shape[0] // (heads * parts), *rest).transpose(0, 1).reshape(t.shape)


def blocked_to_packed(t: torch.Tensor, heads: int, parts: int = 3) -> torch.Tensor:
    """Inverse of :func:`packed_to_blocked` (e.g. to import ``torch.nn.MultiheadAttention.in_proj_weight``)."""
    rest = t.shape[1:]
    return t.reshape(parts, heads, t.shape[0] // (heads * parts), *rest).transpose(0, 1).reshape(t.shape)


def fast_mask_softmax_dropout_func(is_training, heads, inputs, pad_mask, mask_additive, dropout_prob):
    """softmax(inputs + mask) -> dropout. inputs [b*heads, sq, sk]; pad_mask [b, sk] (bool/uint8 = masked, or additive float)."""
    bh, sq, sk = inputs.shape
    b = bh // heads
    x = inputs.view(b, heads, sq, sk)
    if pad_mask is not None and mask_additive:
        x = x + pad_mask.view(b, 1, 1, sk).to(x.dtype)
        p = scaled_softmax(x, 1.0)
    elif pad_mask is not None:
        p = scaled_masked_softmax(x, pad_mask.view(b, 1, 1, sk).expand(b, 1, sq, sk).to(torch.uint8), 1.0)
    else:
        p = scaled_softmax(x, 1.0)
    p = F.dropout(p, dropout_prob, is_training)
    return p.view(bh, sq, sk)


_causal_cache: dict = {}


def _is_causal_mask(attn_mask: torch.Tensor) -> bool:
    """True when the time mask is exactly "mask every key after the query" (one device comparison per distinct mask tensor version)."""
    key = (attn_mask.data_ptr(), tuple(attn_mask.shape), attn_mask._version, attn_mask.dtype)
    hit = _causal_cache.get(key)
    if hit is None:
        tq, tk = attn_mask.shape[-2:]
        tri = torch.ones(tq, tk, dtype=torch.bool, device=attn_mask.device).triu(1)
        hit = bool(tq == tk and attn_mask.dim() == 2 and torch.equal(attn_mask.to(torch.bool) if attn_mask.dtype != torch.bool else attn_mask, tri))
        if len(_causal_cache) > 64:
            _causal_cache.clear()
        _causal_cache[key] = hit
    return hit


def _attention_kernel(q, k, v, heads, scaling, key_bias, causal, dropout):
    """tcgen05 attention (contrib/fmha/kernels.py) on [t, b, heads, hd] views: every (batch, head) pair is one "head" of a single
    batch whose rows are the time steps — the 3-D TMA maps take the row / head strides as they are, so the packed projection output
    is consumed in place and the context comes out as [tq, b, e]."""
    from ..fmha import kernels as K

    tq, b = q.shape[0], q.shape[1]
    hd = q.shape[-1]
    rows = [t.flatten(1, 2) for t in (q, k, v)]                                   # [t, b * heads, hd] (a view for the packed layouts)
    rows = [t if (t.stride(2) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0) else t.contiguous() for t in rows]
    out = K.FmhaFunc.apply(rows[0], rows[1], rows[2], None, None, None, None, 1, causal, float(scaling), key_bias, float(dropout), heads)
    return out.view(tq, b, heads * hd)


def _attention(q, k, v, heads, scaling, key_padding_mask, attn_mask, mask_additive, dropout, training):
    """q [tq, b, heads, hd]; k, v [tk, b, heads, hd] (views) -> [tq, b, e]"""
    tq, b, _, hd = q.shape
    tk = k.shape[0]
    e = heads * hd
    p_drop = float(dropout) if training else 0.0
    from ..fmha import kernels as K

    if K.supported(q, hd) and (attn_mask is None or _is_causal_mask(attn_mask)):
        key_bias = None
        if key_padding_mask is not None:
            kp = key_padding_mask.view(b, tk)
            key_bias = kp.float() if mask_additive else torch.zeros(b, tk, dtype=torch.float32, device=q.device).masked_fill_(kp.to(torch.bool), float("-inf"))
        return _attention_kernel(q, k, v, heads, scaling, key_bias, attn_mask is not None, p_drop)
    q = q.reshape(tq, b * heads, hd).transpose(0, 1)
    k = k.reshape(tk, b * heads, hd).transpose(0, 1)
    v = v.reshape(tk, b * heads, hd).transpose(0, 1)
    scores = torch.bmm(q, k.transpose(1, 2)) * scaling                 # [b*h, tq, tk]
    x = scores.view(b, heads, tq, tk)
    if attn_mask is not None:                                          # time mask [tq, tk] (e.g. causal): True/1 = masked
        m = attn_mask.to(torch.bool).view(1, 1, tq, tk).expand(b, 1, tq, tk).to(torch.uint8)
        p = scaled_masked_softmax(x, m, 1.0)
    elif key_padding_mask is not None and mask_additive:
        p = scaled_softmax(x + key_padding_mask.view(b, 1, 1, tk).to(x.dtype), 1.0)
    elif key_padding_mask is not None:
        m = key_padding_mask.to(torch.bool).view(b, 1, 1, tk).expand(b, 1, tq, tk).to(torch.uint8)
        p = scaled_masked_softmax(x, m, 1.0)
    else:
        p = scaled_softmax(x, 1.0)
    p = F.dropout(p, p_drop, p_drop > 0.0).view(b * heads, tq, tk)
    ctx = torch.bmm(p.to(v.dtype), v)                                   # [b*h, tq, hd]
    return ctx.transpose(0, 1).contiguous().view(tq, b, e)


class _MHABase(nn.Module):
    def _init_norm(self, embed_dim, include_norm_add, impl):
        """State-dict names of the reference: impl="fast" keeps the pre-LN affine as ``lyr_nrm_gamma_weights`` / ``lyr_nrm_beta_weights``,
        impl="default" as a ``lyr_nrm`` FusedLayerNorm sub-module (self_multihead_attn.py:86-96)."""
        self.lyr_nrm = None
        if include_norm_add and impl == "fast":
            self.lyr_nrm_gamma_weights = Parameter(torch.ones(embed_dim))
            self.lyr_nrm_beta_weights = Parameter(torch.zeros(embed_dim))
        else:
            self.register_parameter("lyr_nrm_gamma_weights", None)
            self.register_parameter("lyr_nrm_beta_weights", None)
            if include_norm_add:
                self.lyr_nrm = FusedLayerNorm(embed_dim)

    def _norm(self, query):
        if not self.include_norm_add:
            return query
        if self.lyr_nrm is not None:
            return self.lyr_nrm(query)
        return fused_layer_norm_affine(query, self.lyr_nrm_gamma_weights, self.lyr_nrm_beta_weights, (self.embed_dim,), 1e-5)

    def _reset_norm(self):
        if self.lyr_nrm_gamma_weights is not None:
            nn.init.ones_(self.lyr_nrm_gamma_weights)
            nn.init.zeros_(self.lyr_nrm_beta_weights)
        elif self.lyr_nrm is not None:
            self.lyr_nrm.reset_parameters()

    def _post(self, outputs, query, is_training):
        if self.include_norm_add:
            outputs = F.dropout(outputs, self.dropout, is_training) + query
        return outputs


class SelfMultiheadAttn(_MHABase):
    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=False, include_norm_add=False, impl="fast", separate_qkv_params=False,
                 mask_additive=False):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        assert impl in ("fast", "default"), f"Unsupported impl: {impl} !"
        self.bias, self.include_norm_add, self.impl = bias, include_norm_add, impl
        self.scaling = self.head_dim ** -0.5
        self.separate_qkv_params, self.mask_additive = separate_qkv_params, mask_additive
        if mask_additive:
            assert not include_norm_add, "additive mask not supported with layer norm"
        if separate_qkv_params:
            self.q_weight = Parameter(torch.empty(embed_dim, embed_dim))
            self.k_weight = Parameter(torch.empty(embed_dim, embed_dim))
            self.v_weight = Parameter(torch.empty(embed_dim, embed_dim))
        else:
            self.in_proj_weight = Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.out_proj_weight = Parameter(torch.empty(embed_dim, embed_dim))
        if bias:
            if separate_qkv_params:
                self.q_bias, self.k_bias, self.v_bias = (Parameter(torch.empty(embed_dim)) for _ in range(3))
            else:
                self.in_proj_bias = Parameter(torch.empty(3 * embed_dim))
            self.out_proj_bias = Parameter(torch.empty(embed_dim))
        else:
            for n in (("q_bias", "k_bias", "v_bias") if separate_qkv_params else ("in_proj_bias",)) + ("out_proj_bias",):
                self.register_parameter(n, None)
        self._init_norm(embed_dim, include_norm_add, impl)
        self.reset_parameters()

    def reset_parameters(self):
        if self.separate_qkv_params:
            for w in (self.q_weight, self.k_weight, self.v_weight):
                nn.init.xavier_uniform_(w)
        else:
            nn.init.xavier_uniform_(self.in_proj_weight, gain=math.sqrt(2))
        nn.init.xavier_uniform_(self.out_proj_weight)
        self._reset_norm()
        for n in ("q_bias", "k_bias", "v_bias", "in_proj_bias", "out_proj_bias"):
            b = getattr(self, n, None)
            if b is not None:
                nn.init.constant_(b, 0.0)

    def forward(self, query, key=None, value=None, key_padding_mask=None, need_weights=False, attn_mask=None, is_training=True):
        x = self._norm(query)
        if self.separate_qkv_params:
            w = torch.cat((self.q_weight, self.k_weight, self.v_weight), 0)
            bias = torch.cat((self.q_bias, self.k_bias, self.v_bias), 0) if self.bias else None
        else:
            w, bias = self.in_proj_weight, self.in_proj_bias
        qkv = fused_dense_function(x, w, bias)                          # [t, b, 3e]
        t, b = qkv.shape[0], qkv.shape[1]
        if self.separate_qkv_params:
            q, k, v = (c.view(t, b, self.num_heads, self.head_dim) for c in qkv.chunk(3, dim=-1))
        else:   # packed projection: the reference's per-head interleave [heads, 3, head_dim] (self_multihead_attn_func.py:58-66)
            qkv = qkv.view(t, b, self.num_heads, 3, self.head_dim)
            q, k, v = (qkv[:, :, :, i, :] for i in range(3))
        ctx = _attention(q, k, v, self.num_heads, self.scaling, key_padding_mask, attn_mask, self.mask_additive, self.dropout, is_training)
        out = fused_dense_function(ctx, self.out_proj_weight, self.out_proj_bias)
        return self._post(out, query, is_training), None


class EncdecMultiheadAttn(_MHABase):
    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=False, include_norm_add=False, impl="fast"):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        assert impl in ("fast", "default"), f"Unsupported impl: {impl} !"
        self.bias, self.include_norm_add, self.impl = bias, include_norm_add, impl
        self.scaling = self.head_dim ** -0.5
        self.mask_additive = False
        self.in_proj_weight_q = Parameter(torch.empty(embed_dim, embed_dim))
        self.in_proj_weight_kv = Parameter(torch.empty(2 * embed_dim, embed_dim))
        self.out_proj_weight = Parameter(torch.empty(embed_dim, embed_dim))
        if bias:
            self.in_proj_bias_q = Parameter(torch.empty(embed_dim))
            self.in_proj_bias_kv = Parameter(torch.empty(2 * embed_dim))
            self.out_proj_bias = Parameter(torch.empty(embed_dim))
        else:
            for n in ("in_proj_bias_q", "in_proj_bias_kv", "out_proj_bias"):
                self.register_parameter(n, None)
        self._init_norm(embed_dim, include_norm_add, impl)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.in_proj_weight_q)
        nn.init.xavier_uniform_(self.in_proj_weight_kv, gain=math.sqrt(1.5))
        nn.init.xavier_uniform_(self.out_proj_weight)
        self._reset_norm()
        for n in ("in_proj_bias_q", "in_proj_bias_kv", "out_proj_bias"):
            b = getattr(self, n, None)
            if b is not None:
                nn.init.constant_(b, 0.0)

    def forward(self, query, key, value=None, key_padding_mask=None, need_weights=False, attn_mask=None, is_training=True):
        x = self._norm(query)
        q = fused_dense_function(x, self.in_proj_weight_q, self.in_proj_bias_q)
        kv = fused_dense_function(key, self.in_proj_weight_kv, self.in_proj_bias_kv)
        tk, b = kv.shape[0], kv.shape[1]   # per-head interleave [heads, 2, head_dim] of the reference (encdec_multihead_attn_func.py:86-91)
        kv = kv.view(tk, b, self.num_heads, 2, self.head_dim)
        k, v = (kv[:, :, :, i, :] for i in range(2))
        q = q.view(q.shape[0], b, self.num_heads, self.head_dim)
        ctx = _attention(q, k, v, self.num_heads, self.scaling, key_padding_mask, attn_mask, False, self.dropout, is_training)
        out = fused_dense_function(ctx, self.out_proj_weight, self.out_proj_bias)
        return self._post(out, query, is_training), None
