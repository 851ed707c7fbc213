"""Functional forms of the attention blocks, with the argument lists of the reference's autograd Functions
(apex/contrib/multihead_attn/{self,encdec}_multihead_attn_func.py, fast_*_func.py, *_norm_add_func.py, mask_softmax_dropout_func.py).
The reference hand-writes each backward around its C++ kernels; here every form is the composition the modules use — tcgen05 GEMM projections,
fused LayerNorm, the tcgen05 attention kernels (or fused masked softmax on the generic path) — so autograd derives the backward and there is one implementation of the math.

The reference spreads these over nine files; ``apex_b200.install_as_apex()`` resolves each of those import paths
(``apex.contrib.multihead_attn.self_multihead_attn_func`` ...) to this module."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ...fused_dense import fused_dense_function
from ...normalization.fused_layer_norm import fused_layer_norm_affine
from .multihead_attn import EncdecMultiheadAttn, SelfMultiheadAttn, _attention, fast_mask_softmax_dropout_func  # noqa: F401


def jit_dropout_add(x, residual, prob, is_training):
    """dropout(x) + residual (the reference scripts this with TorchScript, self_multihead_attn.py:24-29)."""
    return F.dropout(x, p=prob, training=is_training) + residual


def _split(lin, heads, parts):
    """[t, b, parts * e] projection with per-head interleaved rows -> ``parts`` views [t, b, heads, head_dim] (no copies)."""
    t, b, pe = lin.shape
    e = pe // parts
    lin = lin.view(t, b, heads, parts, e // heads)
    return [lin[:, :, :, i, :] for i in range(parts)]


def _masks(use_time_mask, mask):
    """(key_padding_mask, attn_mask) from the reference's (use_time_mask, mask) pair."""
    return (None, mask) if use_time_mask else (mask, None)


def _self(use_time_mask, is_training, heads, scale, inputs, w_in, w_out, b_in, b_out, mask, additive, dropout_prob):
    q, k, v = _split(fused_dense_function(inputs, w_in, b_in), heads, 3)
    kpm, am = _masks(use_time_mask, mask)
    ctx = _attention(q, k, v, heads, scale, kpm, am, bool(additive), dropout_prob, is_training)
    return fused_dense_function(ctx, w_out, b_out)


def _encdec(use_time_mask, is_training, heads, scale, inputs_q, inputs_kv, w_q, w_kv, w_out, b_q, b_kv, b_out, mask, dropout_prob):
    q = fused_dense_function(inputs_q, w_q, b_q)
    q = q.view(q.shape[0], q.shape[1], heads, q.shape[2] // heads)
    k, v = _split(fused_dense_function(inputs_kv, w_kv, b_kv), heads, 2)
    kpm, am = _masks(use_time_mask, mask)
    ctx = _attention(q, k, v, heads, scale, kpm, am, False, dropout_prob, is_training)
    return fused_dense_function(ctx, w_out, b_out)


def _scale(t, heads):
    return (t.shape[-1] // heads) ** -0.5


class _Named:
    """``XxxFunc.apply`` spelling of the reference."""


class SelfAttnFunc(_Named):
    @staticmethod
    def apply(use_time_mask, is_training, heads, scale, inputs, input_weights, output_weights, input_biases, output_biases, mask, is_additive_mask,
              dropout_prob):
        return _self(use_time_mask, is_training, heads, scale, inputs, input_weights, output_weights, input_biases, output_biases, mask,
                     is_additive_mask, dropout_prob)


class FastSelfAttnFunc(_Named):
    @staticmethod
    def apply(use_time_mask, is_training, heads, inputs, input_weights, output_weights, input_biases, output_biases, pad_mask, mask_additive,
              dropout_prob):
        return _self(use_time_mask, is_training, heads, _scale(inputs, heads), inputs, input_weights, output_weights, input_biases, output_biases,
                     pad_mask, mask_additive, dropout_prob)


class FastSelfAttnNormAddFunc(_Named):
    @staticmethod
    def apply(use_time_mask, is_training, heads, inputs, lyr_nrm_gamma_weights, lyr_nrm_beta_weights, input_weights, output_weights, pad_mask,
              dropout_prob):
        x = fused_layer_norm_affine(inputs, lyr_nrm_gamma_weights, lyr_nrm_beta_weights, (inputs.shape[-1],), 1e-5)
        out = _self(use_time_mask, is_training, heads, _scale(inputs, heads), x, input_weights, output_weights, None, None, pad_mask, False, dropout_prob)
        return jit_dropout_add(out, inputs, dropout_prob, is_training)


class EncdecAttnFunc(_Named):
    @staticmethod
    def apply(use_time_mask, is_training, heads, scale, inputs_q, inputs_kv, input_weights_q, input_weights_kv, output_weights, input_biases_q,
              input_biases_kv, output_biases, mask, dropout_prob):
        return _encdec(use_time_mask, is_training, heads, scale, inputs_q, inputs_kv, input_weights_q, input_weights_kv, output_weights,
                       input_biases_q, input_biases_kv, output_biases, mask, dropout_prob)


class FastEncdecAttnFunc(_Named):
    @staticmethod
    def apply(use_time_mask, is_training, heads, inputs_q, inputs_kv, input_weights_q, input_weights_kv, output_weights, pad_mask, dropout_prob):
        return _encdec(use_time_mask, is_training, heads, _scale(inputs_q, heads), inputs_q, inputs_kv, input_weights_q, input_weights_kv,
                       output_weights, None, None, None, pad_mask, dropout_prob)


class FastEncdecAttnNormAddFunc(_Named):
    @staticmethod
    def apply(use_time_mask, is_training, heads, inputs_q, inputs_kv, lyr_nrm_gamma_weights, lyr_nrm_beta_weights, input_weights_q,
              input_weights_kv, output_weights, pad_mask, dropout_prob):
        x = fused_layer_norm_affine(inputs_q, lyr_nrm_gamma_weights, lyr_nrm_beta_weights, (inputs_q.shape[-1],), 1e-5)
        out = _encdec(use_time_mask, is_training, heads, _scale(inputs_q, heads), x, inputs_kv, input_weights_q, input_weights_kv, output_weights,
                      None, None, None, pad_mask, dropout_prob)
        return jit_dropout_add(out, inputs_q, dropout_prob, is_training)


class MaskSoftmaxDropout(_Named):
    apply = staticmethod(fast_mask_softmax_dropout_func)


self_attn_func = SelfAttnFunc.apply
fast_self_attn_func = FastSelfAttnFunc.apply
fast_self_attn_norm_add_func = FastSelfAttnNormAddFunc.apply
encdec_attn_func = EncdecAttnFunc.apply
fast_encdec_attn_func = FastEncdecAttnFunc.apply
fast_encdec_attn_norm_add_func = FastEncdecAttnNormAddFunc.apply
