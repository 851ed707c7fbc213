"""Raw ``fast_multihead_attn`` entry points (reference apex/contrib/csrc/multihead_attn/multihead_attn_frontend.cpp:573-607): the 8
forward / backward pairs that hand the intermediate tensors of the attention pipeline to the caller, who passes them back for the
backward. The modules in this package do not go through them (they run the tcgen05 attention kernels, which never materialise the
probabilities); these exist for code that imports the extension by name, and are composed from batched library GEMMs plus this
library's softmax / LayerNorm ops, with the reference's layouts:

* ``input_lin_results`` [T, B, 3E] = per head (q, k, v) interleaved: viewed [T, B*heads, 3, hd]; encdec: q [Tq, B, E], kv [Tk, B, 2E];
* ``softmax_results`` / ``dropout_results`` / ``dropout_mask`` [B*heads, Tq, Tk] (mask uint8, 1 = kept; results already scaled by 1/(1-p));
* ``matmul2_results`` [Tq, B*heads, hd]; ``outputs`` [Tq, B, E]; LayerNorm statistics fp32 [Tq*B].
Masks: ``use_time_mask`` -> uint8 [Tq, Tk], otherwise key padding uint8 [B, Tk] (1 = masked); the additive variants take a
floating [B, Tk] mask that is added to the scores."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _acc(t):
    """fp32 accumulation for the 16-bit dtypes; fp32 / fp64 stay as they are."""
    return t.float() if t.dtype in (torch.float16, torch.bfloat16) else t


def _heads_view(lin, heads, parts):
    T, B, width = lin.shape
    hd = width // (heads * parts)
    v = lin.view(T, B * heads, parts, hd)
    return [v[:, :, i].transpose(0, 1) for i in range(parts)], hd          # each [B*heads, T, hd]


def _masked_scores(scores, heads, use_mask, use_time_mask, pad_mask, additive=False):
    if not use_mask:
        return scores
    if additive:
        BH, Tq, Tk = scores.shape
        return (scores.view(BH // heads, heads, Tq, Tk) + pad_mask.to(scores.dtype)[:, None, None, :]).view(BH, Tq, Tk)
    if use_time_mask:
        return scores.masked_fill(pad_mask.bool()[None], float("-inf"))
    BH, Tq, Tk = scores.shape
    return scores.view(BH // heads, heads, Tq, Tk).masked_fill(pad_mask.bool()[:, None, None, :], float("-inf")).view(BH, Tq, Tk)


def _softmax(scores):
    return torch.softmax(_acc(scores), dim=-1).to(scores.dtype)


def _dropout(x, is_training, p):
    if not is_training or p <= 0.0:
        return x, torch.ones(x.shape, dtype=torch.uint8, device=x.device)
    keep = torch.rand(x.shape, device=x.device) >= p
    return x * keep.to(x.dtype) / (1.0 - p), keep.to(torch.uint8)


def _softmax_bwd(probs, grad):
    pf, gf = _acc(probs), _acc(grad)
    return (pf * (gf - (pf * gf).sum(-1, keepdim=True))).to(probs.dtype)


def _core_fwd(q, k, v, heads, use_mask, use_time_mask, is_training, pad_mask, dropout_prob, additive=False):
    """q [BH, Tq, hd], k / v [BH, Tk, hd] -> (scaled scores before the mask, probabilities, dropped probabilities, keep mask,
    context [Tq, BH, hd])."""
    scale = q.shape[-1] ** -0.5
    bmm1 = torch.bmm(q, k.transpose(1, 2)) * scale
    probs = _softmax(_masked_scores(bmm1, heads, use_mask, use_time_mask, pad_mask, additive))
    dropped, keep = _dropout(probs, is_training, dropout_prob)
    ctx = torch.bmm(dropped, v).transpose(0, 1).contiguous()
    return bmm1, probs, dropped, keep, ctx


def _core_bwd(ctx_grad, q, k, v, probs, dropped, keep, dropout_prob):
    """ctx_grad [Tq, BH, hd] -> dq, dk, dv in the [BH, T, hd] views."""
    g = ctx_grad.transpose(0, 1)
    dv = torch.bmm(dropped.transpose(1, 2), g)
    dp = torch.bmm(g, v.transpose(1, 2)) * keep.to(g.dtype) / (1.0 - dropout_prob)
    ds = _softmax_bwd(probs, dp) * (q.shape[-1] ** -0.5)
    return torch.bmm(ds, k), torch.bmm(ds.transpose(1, 2), q), dv


def _pack(parts_grads, like):
    """[BH, T, hd] gradients of the interleaved parts -> the [T, B, parts*E] layout of the projection output."""
    T, B, width = like.shape
    g = torch.stack([p.transpose(0, 1) for p in parts_grads], dim=2)         # [T, BH, parts, hd]
    return g.reshape(T, B, width)


def _lin_bwd(grad_out, x, w):
    g2, x2 = grad_out.reshape(-1, grad_out.shape[-1]), x.reshape(-1, x.shape[-1])
    return (g2 @ w).view_as(x), g2.t() @ x2


def _layer_norm_fwd(x, gamma, beta, eps=1e-5):
    xf = _acc(x)
    mean = xf.mean(-1)
    invvar = torch.rsqrt(xf.var(-1, unbiased=False) + eps)
    y = ((xf - mean[..., None]) * invvar[..., None] * _acc(gamma) + _acc(beta)).to(x.dtype)
    return y, mean.reshape(-1), invvar.reshape(-1)


def _layer_norm_bwd(gy, x, mean, invvar, gamma):
    xf, gf = _acc(x), _acc(gy)
    xhat = (xf - mean.view(*x.shape[:-1], 1)) * invvar.view(*x.shape[:-1], 1)
    gg = gf * _acc(gamma)
    dx = (gg - gg.mean(-1, keepdim=True) - xhat * (gg * xhat).mean(-1, keepdim=True)) * invvar.view(*x.shape[:-1], 1)
    red = tuple(range(x.dim() - 1))
    return dx.to(x.dtype), (gf * xhat).sum(red).to(gamma.dtype), gf.sum(red).to(gamma.dtype)


# ---- softmax + dropout on scores the caller computed ------------------------------------------------------------------------------
def mask_softmax_dropout_forward(use_mask, is_training, heads, input, pad_mask, dropout_prob):
    probs = _softmax(_masked_scores(input, heads, use_mask, False, pad_mask))
    dropped, keep = _dropout(probs, is_training, dropout_prob)
    return [dropped, keep, probs]


def mask_softmax_dropout_backward(use_mask, heads, output_grads, softmax_results, dropout_mask, padding_mask, dropout_prob):
    return _softmax_bwd(softmax_results, output_grads * dropout_mask.to(output_grads.dtype) / (1.0 - dropout_prob))


def additive_mask_softmax_dropout_forward(use_mask, is_training, heads, input, pad_mask, dropout_prob):
    probs = _softmax(_masked_scores(input, heads, use_mask, False, pad_mask, additive=True))
    dropped, keep = _dropout(probs, is_training, dropout_prob)
    return [dropped, keep, probs]


def additive_mask_softmax_dropout_backward(use_mask, heads, output_grads, softmax_results, dropout_mask, dropout_prob):
    return _softmax_bwd(softmax_results, output_grads * dropout_mask.to(output_grads.dtype) / (1.0 - dropout_prob))


# ---- self attention -----------------------------------------------------------------------------------------------------------------
def _self_fwd(use_mask, use_time_mask, is_training, heads, inputs, w_in, w_out, b_in, b_out, pad_mask, dropout_prob, additive):
    lin = F.linear(inputs, w_in, b_in)
    (q, k, v), _ = _heads_view(lin, heads, 3)
    bmm1, probs, dropped, keep, ctx = _core_fwd(q, k, v, heads, use_mask, use_time_mask, is_training, pad_mask, dropout_prob, additive)
    out = F.linear(ctx.view(inputs.shape[0], inputs.shape[1], -1), w_out, b_out)
    return lin, bmm1, probs, dropped, keep, ctx, out


def _self_bwd(heads, output_grads, matmul2_results, dropout_results, softmax_results, input_lin_results, inputs, w_in, w_out,
              dropout_mask, dropout_prob):
    (q, k, v), _ = _heads_view(input_lin_results, heads, 3)
    T, B, E = inputs.shape
    ctx_grad, w_out_grad = _lin_bwd(output_grads, matmul2_results.view(T, B, -1), w_out)
    dq, dk, dv = _core_bwd(ctx_grad.view_as(matmul2_results), q, k, v, softmax_results, dropout_results, dropout_mask, dropout_prob)
    lin_grad = _pack([dq, dk, dv], input_lin_results)
    in_grad, w_in_grad = _lin_bwd(lin_grad, inputs, w_in)
    return in_grad, w_in_grad, w_out_grad, lin_grad


def self_attn_forward(use_mask, use_time_mask, is_training, heads, inputs, input_weights, output_weights, pad_mask, dropout_prob):
    lin, _, probs, dropped, keep, ctx, out = _self_fwd(use_mask, use_time_mask, is_training, heads, inputs, input_weights, output_weights,
                                                       None, None, pad_mask, dropout_prob, False)
    return [lin, probs, dropped, keep, ctx, out]


def self_attn_backward(heads, output_grads, matmul2_results, dropout_results, softmax_results, input_lin_results, inputs, input_weights,
                       output_weights, dropout_mask, dropout_prob):
    return list(_self_bwd(int(heads), output_grads, matmul2_results, dropout_results, softmax_results, input_lin_results, inputs,
                          input_weights, output_weights, dropout_mask, float(dropout_prob))[:3])


def self_attn_bias_forward(use_mask, use_time_mask, is_training, heads, inputs, input_weights, output_weights, input_biases,
                           output_biases, pad_mask, dropout_prob):
    lin, _, probs, dropped, keep, ctx, out = _self_fwd(use_mask, use_time_mask, is_training, heads, inputs, input_weights, output_weights,
                                                       input_biases, output_biases, pad_mask, dropout_prob, False)
    return [lin, probs, dropped, keep, ctx, out]


def self_attn_bias_backward(heads, output_grads, matmul2_results, dropout_results, softmax_results, input_lin_results, inputs,
                            input_weights, output_weights, dropout_mask, dropout_prob):
    in_grad, w_in_grad, w_out_grad, lin_grad = _self_bwd(int(heads), output_grads, matmul2_results, dropout_results, softmax_results,
                                                         input_lin_results, inputs, input_weights, output_weights, dropout_mask,
                                                         float(dropout_prob))
    return [in_grad, w_in_grad, w_out_grad, lin_grad.sum((0, 1)), output_grads.sum((0, 1))]


def self_attn_bias_additive_mask_forward(use_mask, use_time_mask, is_training, heads, inputs, input_weights, output_weights, input_biases,
                                         output_biases, pad_mask, dropout_prob):
    lin, bmm1, _, dropped, keep, ctx, out = _self_fwd(use_mask, use_time_mask, is_training, heads, inputs, input_weights, output_weights,
                                                      input_biases, output_biases, pad_mask, dropout_prob, True)
    return [lin, bmm1, dropped, keep, ctx, out]


def self_attn_bias_additive_mask_backward(heads, output_grads, matmul2_results, dropout_results, bmm1_results, pad_mask,
                                          input_lin_results, inputs, input_weights, output_weights, dropout_mask, dropout_prob):
    # the probabilities are not passed back: recomputed from the saved scores and the mask, as the reference's backward does
    heads = int(heads)
    use_mask = pad_mask is not None and pad_mask.numel() > 0
    probs = _softmax(_masked_scores(bmm1_results, heads, use_mask, False, pad_mask, additive=True))
    in_grad, w_in_grad, w_out_grad, lin_grad = _self_bwd(heads, output_grads, matmul2_results, dropout_results, probs, input_lin_results,
                                                         inputs, input_weights, output_weights, dropout_mask, float(dropout_prob))
    return [in_grad, w_in_grad, w_out_grad, lin_grad.sum((0, 1)), output_grads.sum((0, 1))]


def self_attn_norm_add_forward(use_mask, use_time_mask, is_training, heads, inputs, lyr_nrm_gamma_weights, lyr_nrm_beta_weights,
                               input_weights, output_weights, pad_mask, dropout_prob):
    normed, mean, invvar = _layer_norm_fwd(inputs, lyr_nrm_gamma_weights, lyr_nrm_beta_weights)
    lin, _, probs, dropped, keep, ctx, out_lin = _self_fwd(use_mask, use_time_mask, is_training, heads, normed, input_weights,
                                                           output_weights, None, None, pad_mask, dropout_prob, False)
    out_dropped, add_keep = _dropout(out_lin, is_training, dropout_prob)
    return [normed, mean, invvar, lin, probs, dropped, keep, ctx, add_keep, out_dropped + inputs]


def self_attn_norm_add_backward(heads, output_grads, matmul2_results, dropout_results, softmax_results, input_lin_results,
                                lyr_nrm_results, lyr_nrm_mean, lyr_nrm_invvar, inputs, lyr_nrm_gamma_weights, lyr_nrm_beta_weights,
                                input_weights, output_weights, dropout_mask, dropout_add_mask, dropout_prob):
    p = float(dropout_prob)
    out_lin_grad = output_grads * dropout_add_mask.to(output_grads.dtype) / (1.0 - p)
    normed_grad, w_in_grad, w_out_grad, _ = _self_bwd(int(heads), out_lin_grad, matmul2_results, dropout_results, softmax_results,
                                                      input_lin_results, lyr_nrm_results, input_weights, output_weights, dropout_mask, p)
    dx, dgamma, dbeta = _layer_norm_bwd(normed_grad, inputs, lyr_nrm_mean, lyr_nrm_invvar, lyr_nrm_gamma_weights)
    return [dx + output_grads, dgamma, dbeta, w_in_grad, w_out_grad]


# ---- encoder-decoder attention ------------------------------------------------------------------------------------------------------
def _encdec_fwd(use_mask, use_time_mask, is_training, heads, inputs_q, inputs_kv, w_q, w_kv, w_out, pad_mask, dropout_prob):
    lin_q, lin_kv = F.linear(inputs_q, w_q), F.linear(inputs_kv, w_kv)
    (q,), _ = _heads_view(lin_q, heads, 1)
    (k, v), _ = _heads_view(lin_kv, heads, 2)
    _, probs, dropped, keep, ctx = _core_fwd(q, k, v, heads, use_mask, use_time_mask, is_training, pad_mask, dropout_prob)
    out = F.linear(ctx.view(inputs_q.shape[0], inputs_q.shape[1], -1), w_out)
    return lin_q, lin_kv, probs, dropped, keep, ctx, out


def _encdec_bwd(heads, output_grads, matmul2_results, dropout_results, softmax_results, lin_q, lin_kv, inputs_q, inputs_kv, w_q, w_kv,
                w_out, dropout_mask, dropout_prob):
    (q,), _ = _heads_view(lin_q, heads, 1)
    (k, v), _ = _heads_view(lin_kv, heads, 2)
    Tq, B, _ = inputs_q.shape
    ctx_grad, w_out_grad = _lin_bwd(output_grads, matmul2_results.view(Tq, B, -1), w_out)
    dq, dk, dv = _core_bwd(ctx_grad.view_as(matmul2_results), q, k, v, softmax_results, dropout_results, dropout_mask, dropout_prob)
    q_grad, w_q_grad = _lin_bwd(_pack([dq], lin_q), inputs_q, w_q)
    kv_grad, w_kv_grad = _lin_bwd(_pack([dk, dv], lin_kv), inputs_kv, w_kv)
    return q_grad, kv_grad, w_q_grad, w_kv_grad, w_out_grad


def encdec_multihead_attn_forward(use_mask, use_time_mask, is_training, heads, inputs_q, inputs_kv, input_weights_q, input_weights_kv,
                                  output_weights, pad_mask, dropout_prob):
    return list(_encdec_fwd(use_mask, use_time_mask, is_training, heads, inputs_q, inputs_kv, input_weights_q, input_weights_kv,
                            output_weights, pad_mask, dropout_prob))


def encdec_multihead_attn_backward(heads, output_grads, matmul2_results, dropout_results, softmax_results, input_lin_q_results,
                                   input_lin_kv_results, inputs_q, inputs_kv, input_weights_q, input_weights_kv, output_weights,
                                   dropout_mask, dropout_prob):
    return list(_encdec_bwd(int(heads), output_grads, matmul2_results, dropout_results, softmax_results, input_lin_q_results,
                            input_lin_kv_results, inputs_q, inputs_kv, input_weights_q, input_weights_kv, output_weights, dropout_mask,
                            float(dropout_prob)))


def encdec_multihead_attn_norm_add_forward(use_mask, use_time_mask, is_training, heads, inputs_q, inputs_kv, lyr_nrm_gamma_weights,
                                           lyr_nrm_beta_weights, input_weights_q, input_weights_kv, output_weights, pad_mask,
                                           dropout_prob):
    normed, mean, invvar = _layer_norm_fwd(inputs_q, lyr_nrm_gamma_weights, lyr_nrm_beta_weights)
    lin_q, lin_kv, probs, dropped, keep, ctx, out_lin = _encdec_fwd(use_mask, use_time_mask, is_training, heads, normed, inputs_kv,
                                                                    input_weights_q, input_weights_kv, output_weights, pad_mask,
                                                                    dropout_prob)
    out_dropped, add_keep = _dropout(out_lin, is_training, dropout_prob)
    return [normed, mean, invvar, lin_q, lin_kv, probs, dropped, keep, ctx, add_keep, out_dropped + inputs_q]


def encdec_multihead_attn_norm_add_backward(heads, output_grads, matmul2_results, dropout_results, softmax_results, input_lin_q_results,
                                            input_lin_kv_results, lyr_nrm_results, lyr_nrm_mean, lyr_nrm_invvar, inputs_q, inputs_kv,
                                            lyr_nrm_gamma_weights, lyr_nrm_beta_weights, input_weights_q, input_weights_kv,
                                            output_weights, dropout_mask, dropout_add_mask, dropout_prob):
    p = float(dropout_prob)
    out_lin_grad = output_grads * dropout_add_mask.to(output_grads.dtype) / (1.0 - p)
    normed_grad, kv_grad, w_q_grad, w_kv_grad, w_out_grad = _encdec_bwd(int(heads), out_lin_grad, matmul2_results, dropout_results,
                                                                        softmax_results, input_lin_q_results, input_lin_kv_results,
                                                                        lyr_nrm_results, inputs_kv, input_weights_q, input_weights_kv,
                                                                        output_weights, dropout_mask, p)
    dx, dgamma, dbeta = _layer_norm_bwd(normed_grad, inputs_q, lyr_nrm_mean, lyr_nrm_invvar, lyr_nrm_gamma_weights)
    return [dx + output_grads, kv_grad, dgamma, dbeta, w_q_grad, w_kv_grad, w_out_grad]


ENTRY_POINTS = tuple(n for n in dir() if n.endswith(("_forward", "_backward")))
