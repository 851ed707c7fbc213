"""Extension-name compatibility: downstream code (Megatron-LM and friends) imports the reference's compiled extensions by name and
calls their raw entry points (``scaled_upper_triang_masked_softmax_cuda.forward``, ``fused_layer_norm_cuda.forward_affine``,
``fused_weight_gradient_mlp_cuda.wgrad_gemm_accum_fp32`` ...). ``install()`` registers modules with those names in ``sys.modules``
whose functions run this library's kernels with the reference's argument order and return conventions
(SURVEY.md §2.12 lists the extensions; signatures from csrc/*.cpp of the reference).
"""
from __future__ import annotations

import sys
import types

import torch


def _mod(name: str, **fns) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__doc__ = f"apex_b200 implementation of the reference extension `{name}`"
    for k, v in fns.items():
        setattr(m, k, v)
    return m


def _softmax_mods():
    from .transformer.functional import fused_softmax as S

    def get_batch_per_block(sq, sk, b, np):  # launch-shape helper of the reference kernels; any value is legal here
        return 1

    return {
        "scaled_softmax_cuda": _mod("scaled_softmax_cuda", forward=lambda x, scale: S._fwd(x, None, scale, 0),
                                    backward=lambda dy, y, scale: S._bwd(dy, y, scale)),
        "scaled_masked_softmax_cuda": _mod("scaled_masked_softmax_cuda", forward=lambda x, mask, scale: S._fwd(x, mask, scale, 1),
                                           backward=lambda dy, y, scale: S._bwd(dy, y, scale), get_batch_per_block=get_batch_per_block),
        "generic_scaled_masked_softmax_cuda": _mod("generic_scaled_masked_softmax_cuda", forward=lambda x, mask, scale: S._fwd(x, mask, scale, 1),
                                                   backward=lambda dy, y, scale: S._bwd(dy, y, scale)),
        "scaled_upper_triang_masked_softmax_cuda": _mod("scaled_upper_triang_masked_softmax_cuda", forward=lambda x, scale: S._fwd(x, None, scale, 2),
                                                        backward=lambda dy, y, scale: S._bwd(dy, y, scale)),
    }


def _layer_norm_mod():
    from .ops import norm as N

    def _cpu(x, shape, w, b, eps, rms):
        from .normalization.custom_ops import reference_fwd

        y, mean, invvar = reference_fwd(x, shape, w, b, eps, rms, w.dtype if w is not None else x.dtype)
        return y, mean, invvar

    def _fwd(x, shape, w, b, eps, rms, mixed=False):
        shape = tuple(shape)
        if not x.is_cuda:
            return _cpu(x, shape, w, b, eps, rms)
        return N.norm_fwd(x, shape, w, b, eps, rms, w.dtype if (mixed and w is not None) else None)

    def _bwd(dy, mean, invvar, saved, shape, w, b, eps, rms, memory_efficient):
        if not dy.is_cuda:
            from .normalization.custom_ops import reference_bwd

            return reference_bwd(dy, saved, mean, invvar, tuple(shape), w, b, rms, memory_efficient, saved.dtype if not memory_efficient else dy.dtype)
        return N.norm_bwd(dy, saved, mean, invvar, tuple(shape), w, b, eps, rms, memory_efficient, saved.dtype if not memory_efficient else dy.dtype)

    return _mod(
        "fused_layer_norm_cuda",
        forward=lambda x, shape, eps: _fwd(x, shape, None, None, eps, False),
        forward_affine=lambda x, shape, w, b, eps: _fwd(x, shape, w, b, eps, False),
        forward_affine_mixed_dtypes=lambda x, shape, w, b, eps: _fwd(x, shape, w, b, eps, False, True),
        backward=lambda dy, mean, invvar, saved, shape, eps, memory_efficient=False: _bwd(dy, mean, invvar, saved, shape, None, None, eps, False, memory_efficient)[0],
        backward_affine=lambda dy, mean, invvar, saved, shape, w, b, eps, memory_efficient=False: _bwd(dy, mean, invvar, saved, shape, w, b, eps, False, memory_efficient),
        rms_forward=lambda x, shape, eps: _fwd(x, shape, None, None, eps, True)[::2],
        rms_forward_affine=lambda x, shape, w, eps: _fwd(x, shape, w, None, eps, True)[::2],
        rms_forward_affine_mixed_dtypes=lambda x, shape, w, eps: _fwd(x, shape, w, None, eps, True, True)[::2],
        rms_backward=lambda dy, invvar, saved, shape, eps, memory_efficient=False: _bwd(dy, None, invvar, saved, shape, None, None, eps, True, memory_efficient)[0],
        rms_backward_affine=lambda dy, invvar, saved, shape, w, eps, memory_efficient=False: _bwd(dy, None, invvar, saved, shape, w, None, eps, True, memory_efficient)[:2],
    )


def _rope_mod():
    from .transformer.functional import fused_rope as R

    def _2d(t, cos_h, sin_h, cos_w, sin_w, is_bwd):
        """The raw extension takes [b, img_h, img_w, heads, d] (fused_rotary_positional_embedding.cpp:137-152) and returns the same shape."""
        b, ih, iw, h, d = t.shape
        return R._2d(t.reshape(b, ih * iw, h, d), ih, iw, cos_h, sin_h, cos_w, sin_w, is_bwd).view(b, ih, iw, h, d)

    return _mod(
        "fused_rotary_positional_embedding",
        forward=lambda t, freqs, transpose_output=False: R._sbhd(t, freqs, None, None, transpose_output, False),
        backward=lambda g, freqs, transpose_output=False: R._sbhd(g, freqs, None, None, transpose_output, True),
        forward_cached=lambda t, cos, sin, transpose_output=False: R._sbhd(t, None, cos, sin, transpose_output, False),
        backward_cached=lambda g, cos, sin, transpose_output=False: R._sbhd(g, None, cos, sin, transpose_output, True),
        forward_thd=lambda t, cu_seqlens, freqs: R._thd(t, cu_seqlens, freqs, False),
        backward_thd=lambda g, cu_seqlens, freqs: R._thd(g, cu_seqlens, freqs, True),
        forward_2d=lambda t, cos_h, sin_h, cos_w, sin_w: _2d(t, cos_h, sin_h, cos_w, sin_w, False),
        backward_2d=lambda g, cos_h, sin_h, cos_w, sin_w: _2d(g, cos_h, sin_h, cos_w, sin_w, True),
    )


def _xentropy_mod():
    from .contrib.xentropy.softmax_xentropy import SoftmaxCrossEntropyLoss as X

    def forward(logits, labels, smoothing, half_to_float):
        ctx = _Ctx()
        losses = X.forward(ctx, logits, labels, smoothing, -1, half_to_float)  # padding is applied by the python layer of the reference
        return losses, ctx.saved_tensors[1]

    def backward(grad_loss, logits, max_log_sum_exp, labels, smoothing):
        ctx = _Ctx()
        ctx.saved_tensors = (logits.contiguous(), max_log_sum_exp, labels.contiguous().view(-1).to(torch.int64))
        ctx.smoothing, ctx.padding_idx = smoothing, -1
        return X.backward(ctx, grad_loss)[0]

    return _mod("xentropy_cuda", forward=forward, backward=backward)


def _dense_mods():
    from .ops import gemm as G

    def linear_bias_forward(input, weight, bias):
        return G.linear_fwd(input.contiguous(), weight.contiguous(), bias)

    def linear_bias_backward(input, weight, d_output):
        dy = d_output.contiguous()
        return [G.linear_dgrad(dy, weight.contiguous()), G.linear_wgrad(dy, input.contiguous()), G.colsum(dy)]

    def linear_gelu_linear_forward(input, weight1, bias1, weight2, bias2):
        x = input.contiguous()
        gelu_in = torch.empty(x.shape[0], weight1.shape[0], dtype=x.dtype, device=x.device)
        output1 = G.linear_fwd(x, weight1.contiguous(), bias1, epi=G.EPI_BIAS_GELU, aux=gelu_in)
        return [output1, G.linear_fwd(output1, weight2.contiguous(), bias2), gelu_in]

    def linear_gelu_linear_backward(input, gelu_in, output1, weight1, weight2, d_output2):
        dy = d_output2.contiguous()
        d_gelu_in = G.linear_dgrad(dy, weight2.contiguous(), dgelu_aux=gelu_in)
        return [G.linear_dgrad(d_gelu_in, weight1.contiguous()), G.linear_wgrad(d_gelu_in, input.contiguous()), G.colsum(d_gelu_in),
                G.linear_wgrad(dy, output1), G.colsum(dy)]

    def _split(use_bias, inputs):
        n = (len(inputs) - 1) // (2 if use_bias else 1)
        return n, inputs[1:1 + n], (inputs[1 + n:1 + 2 * n] if use_bias else [None] * n)

    def mlp_forward(use_bias, activation, inputs):
        """-> [output, hidden_1, ..., hidden_{n-1}]  (the reference returns [output, reserved_space]; pass the list back to backward)."""
        n, ws, bs = _split(use_bias, list(inputs))
        acts = [inputs[0].contiguous()]
        for w, b in zip(ws, bs):
            epi = {0: (G.EPI_NONE, G.EPI_BIAS), 1: (G.EPI_RELU, G.EPI_BIAS_RELU), 2: (G.EPI_SIGMOID, G.EPI_BIAS_SIGMOID)}[int(activation)][b is not None]
            acts.append(G.linear_fwd(acts[-1], w.contiguous(), b, epi=epi))
        return [acts[-1], *acts[1:-1]]

    def mlp_backward(use_bias, activation, grad_o, fprop_outputs, inputs):
        n, ws, _ = _split(use_bias, list(inputs))
        acts = [inputs[0].contiguous(), *fprop_outputs[1:], fprop_outputs[0]]
        dy = grad_o.contiguous()
        dws, dbs = [None] * n, [None] * n
        for i in range(n - 1, -1, -1):
            y = acts[i + 1]
            if activation == 1:
                dy = dy * (y > 0).to(dy.dtype)
            elif activation == 2:
                dy = (dy.float() * (y.float() * (1 - y.float()))).to(dy.dtype)
            dws[i] = G.linear_wgrad(dy, acts[i])
            if use_bias:
                dbs[i] = G.colsum(dy)
            dy = G.linear_dgrad(dy, ws[i].contiguous())
        return [dy, *dws, *(dbs if use_bias else [])]

    return {
        "fused_dense_cuda": _mod("fused_dense_cuda", linear_bias_forward=linear_bias_forward, linear_bias_backward=linear_bias_backward,
                                 linear_gelu_linear_forward=linear_gelu_linear_forward, linear_gelu_linear_backward=linear_gelu_linear_backward),
        "mlp_cuda": _mod("mlp_cuda", forward=mlp_forward, backward=mlp_backward),
    }


def _fast_layer_norm_mod(ln):
    def ln_fwd(x, gamma, beta, epsilon):
        """-> [z, mu, rsigma]  (apex/contrib/csrc/layer_norm/ln_api.cpp:83)"""
        return list(ln.forward_affine_mixed_dtypes(x, (x.shape[-1],), gamma, beta, epsilon))

    def ln_bwd(dz, x_or_z, mu, rsigma, gamma, beta=None, memory_efficient=False):
        """-> [dx, dgamma, dbeta, dgamma_part, dbeta_part]; the two partial buffers of the reference are internal scratch, returned empty."""
        dx, dg, db = ln.backward_affine(dz, mu, rsigma, x_or_z, (x_or_z.shape[-1],), gamma, beta if beta is not None else torch.zeros_like(gamma),
                                        0.0, memory_efficient)
        e = dg.new_empty(0)
        return [dx, dg, db, e, e]

    return _mod("fast_layer_norm", ln_fwd=ln_fwd, ln_bwd=ln_bwd)


def _small_contrib_mods():
    import importlib

    from . import _lib

    # the packages re-export a FUNCTION under the sub-module's own name (focal_loss, index_mul_2d): import the modules explicitly
    FL = importlib.import_module(__package__ + ".contrib.focal_loss.focal_loss")
    IM = importlib.import_module(__package__ + ".contrib.index_mul_2d.index_mul_2d")

    def focal_forward(cls_output, cls_targets_at_level, num_positives_sum, num_real_classes, alpha, gamma, smoothing_factor):
        """-> [loss, partial_grad]"""
        if not cls_output.is_cuda:
            with torch.enable_grad():
                x = cls_output.detach().requires_grad_()
                loss = FL._ref(x, cls_targets_at_level, num_positives_sum, num_real_classes, alpha, gamma, smoothing_factor)
                (pgrad,) = torch.autograd.grad(loss, x)
            return [loss.detach(), pgrad]
        ctx = _Ctx()
        loss = FL.FocalLoss.forward(ctx, cls_output, cls_targets_at_level, num_positives_sum, num_real_classes, alpha, gamma, smoothing_factor)
        return [loss, ctx.saved_tensors[0]]

    def focal_backward(grad_output, partial_grad, num_positives_sum):
        if not partial_grad.is_cuda:
            return partial_grad * grad_output
        ctx = _Ctx()
        ctx.saved_tensors = (partial_grad, num_positives_sum.float().reshape(1))
        return FL.FocalLoss.backward(ctx, grad_output)[0]

    def im_forward(out, in1, in2, idx1):
        if not IM._native(in1):
            out.copy_(in1.index_select(0, idx1) * in2)
            return
        _lib.fn("ab_index_mul_2d_fwd")(in1.data_ptr(), in2.data_ptr(), idx1.data_ptr(), out.data_ptr(), in2.shape[0], in2.shape[1], _lib.dt(in1),
                                       _lib.stream_ptr(in1.device))

    def im_backward(grad_in1, grad_in2, grad_out, in1, in2, idx1):
        """grad_in1 must arrive zero-filled (it is accumulated into), as in the reference."""
        if not IM._native(in1):
            grad_in1.index_add_(0, idx1, (grad_out * in2).to(grad_in1.dtype))
            grad_in2.copy_(grad_out * in1.index_select(0, idx1))
            return
        g1, g2 = IM._IndexMul2dBackward.forward(_Ctx(), in1, in2, idx1, grad_out.contiguous())
        grad_in1.add_(g1)
        grad_in2.copy_(g2)

    def im_backward_backward(grad_grad_out, grad_in1, grad_in2, grad_out, grad_grad_in1, grad_grad_in2, in1, in2, idx1):
        gg1 = grad_grad_in1.index_select(0, idx1)
        grad_in1.index_add_(0, idx1, (grad_grad_in2 * grad_out).to(grad_in1.dtype))
        grad_in2.copy_(gg1 * grad_out)
        grad_grad_out.copy_(gg1 * in2 + grad_grad_in2 * in1.index_select(0, idx1))

    return {
        "focal_loss_cuda": _mod("focal_loss_cuda", forward=focal_forward, backward=focal_backward),
        "fused_index_mul_2d": _mod("fused_index_mul_2d", float_forward=im_forward, float_backward=im_backward, float_backward_backward=im_backward_backward,
                                   half_forward=im_forward, half_backward=im_backward, half_backward_backward=im_backward_backward),
    }


def _distopt_mods():
    from . import _lib
    from .ops import amp_C
    from .ops import reference as ref

    def _cuda(lists):
        return len(lists[0]) > 0 and lists[0][0].is_cuda

    def multi_tensor_fused_adam(chunk_size, noop_flag, tensor_lists, grad_scale, lr, beta1, beta2, eps, step, mode, bias_correction, weight_decay):
        """[p_in, m, v, g, p_out]  (apex/contrib/csrc/optimizers/multi_tensor_distopt_adam.cpp:3)"""
        if not tensor_lists or not tensor_lists[0]:
            return
        if not _cuda(tensor_lists):
            for p_in, m, v, g, p_out in zip(*tensor_lists):
                ref.dist_adam(p_in, m, v, g, p_out, grad_scale, lr, beta1, beta2, eps, step, mode, bias_correction, weight_decay)
            return
        tb = amp_C.TensorTable(tensor_lists, chunk_size)
        d = tb.dtypes
        _lib.fn("ab_mt_dist_adam")(*tb.head(), d[0], d[3], d[4], grad_scale.data_ptr(), float(lr), float(beta1), float(beta2), float(eps), int(step),
                                   int(mode), int(bias_correction), float(weight_decay), 0, None, None, None, _lib.stream_ptr(tb.device))

    def multi_tensor_fused_adam_capturable(chunk_size, noop_flag, tensor_lists, grad_scale, lr, beta1, beta2, eps, step, mode, bias_correction,
                                           weight_decay):
        """lr / step are device tensors; nothing is written when ``noop_flag`` is set."""
        if not tensor_lists or not tensor_lists[0]:
            return
        if not _cuda(tensor_lists):
            if int(noop_flag.item()) == 1:
                return
            for p_in, m, v, g, p_out in zip(*tensor_lists):
                ref.dist_adam(p_in, m, v, g, p_out, grad_scale, float(lr), beta1, beta2, eps, int(step), mode, bias_correction, weight_decay)
            return
        tb = amp_C.TensorTable(tensor_lists, chunk_size)
        d = tb.dtypes
        step_i = step if step.dtype == torch.int32 else step.to(torch.int32)
        lr_f = lr if lr.dtype == torch.float32 else lr.float()
        _lib.fn("ab_mt_dist_adam")(*tb.head(), d[0], d[3], d[4], grad_scale.data_ptr(), 0.0, float(beta1), float(beta2), float(eps), 0, int(mode),
                                   int(bias_correction), float(weight_decay), 1, lr_f.data_ptr(), step_i.data_ptr(), _lib.ptr(noop_flag),
                                   _lib.stream_ptr(tb.device))

    def multi_tensor_fused_adam_with_param_remainders(chunk_size, noop_flag, tensor_lists, grad_scale, lr, beta1, beta2, eps, step, mode,
                                                      bias_correction, weight_decay):
        """[p_in (bf16 bits as int16), p_remainder (int16), m, v, g, p_out (bf16)]: fp32 master = (p_in << 16) + remainder."""
        if not tensor_lists or not tensor_lists[0]:
            return
        if not _cuda(tensor_lists):
            for p_hi, p_lo, m, v, g, p_out in zip(*tensor_lists):
                hi, lo = p_hi.view(torch.int16).to(torch.int32), p_lo.to(torch.int32)
                master = ((hi << 16) + lo).view(torch.float32).clone()   # remainder is signed: hi was rounded to nearest
                ref.dist_adam(master, m, v, g, None, grad_scale, lr, beta1, beta2, eps, step, mode, bias_correction, weight_decay)
                bits = master.view(torch.int32)
                new_lo = ((bits & 0xFFFF) ^ 0x8000) - 0x8000             # sign-extended low half
                new_hi = (bits - new_lo) >> 16
                p_lo.copy_(new_lo.to(torch.int16))
                p_out.view(torch.int16).copy_(new_hi.to(torch.int16))    # p_in is read-only (it normally aliases p_out)
            return
        lists = [[t.view(torch.int16) for t in tensor_lists[0]], list(tensor_lists[1]), list(tensor_lists[2]), list(tensor_lists[3]),
                 list(tensor_lists[4]), [t.view(torch.int16) for t in tensor_lists[5]]]
        tb = amp_C.TensorTable(lists, chunk_size)
        _lib.fn("ab_mt_dist_adam_remainders")(*tb.head(), tb.dtypes[4], grad_scale.data_ptr(), float(lr), float(beta1), float(beta2), float(eps),
                                              int(step), int(mode), int(bias_correction), float(weight_decay), _lib.stream_ptr(tb.device))

    def multi_tensor_lamb_compute_update_term(chunk_size, noop_flag, tensor_lists, per_tensor_beta1, per_tensor_beta2, per_tensor_beta3,
                                              per_tensor_bias_correction, step, per_tensor_epsilon, mode, per_tensor_decay, global_scale,
                                              global_grad_norm, max_grad_norm):
        """[g, p, m, v, u]: Adam-style update term with per-tensor hyper-parameters held in device tensors
        (multi_tensor_distopt_lamb_kernel.cu:98-274 of the reference). Off this library's hot path (DistributedFusedLAMB uses the fused
        stage kernels); expressed with torch ops so device scalars never reach the host."""
        if int(noop_flag.item()) == 1:
            return
        gs = global_scale.float().reshape(())
        cs = gs
        if max_grad_norm > 0:
            c = max_grad_norm / (global_grad_norm.float().reshape(()) / gs + 1e-6)
            cs = gs / torch.clamp(c, max=1.0)
        stepf = step.float().reshape(())
        for i, (g, p, m, v, u) in enumerate(zip(*tensor_lists)):
            b1, b2, eps, decay = per_tensor_beta1[i], per_tensor_beta2[i], per_tensor_epsilon[i], per_tensor_decay[i]
            bc = per_tensor_bias_correction[i] == 1
            c1 = torch.where(bc, 1 - b1 ** stepf, torch.ones_like(b1))
            c2 = torch.where(bc, 1 - b2 ** stepf, torch.ones_like(b2))
            sg = g.float() / cs
            pf = p.float()
            if mode == 0:
                sg = sg + decay * pf
            mf = m.float() * b1 + (1 - b1) * sg
            vf = v.float() * b2 + (1 - b2) * sg * sg
            upd = (mf / c1) / ((vf / c2).sqrt() + eps)
            if mode != 0:
                upd = upd + decay * pf
            m.copy_(mf)
            v.copy_(vf)
            u.copy_(upd)

    def multi_tensor_lamb_update_weights(chunk_size, noop_flag, tensor_lists, per_tensor_param_norm, per_tensor_update_norm, update_norm_offset,
                                         learning_rate, per_tensor_decay, global_grad_norm, use_nvlamb):
        """[u, p, p_copy]: p -= lr * trust_ratio * u, trust ratio only where decay != 0 unless ``use_nvlamb`` (kernel :276-360)."""
        if int(noop_flag.item()) == 1:
            return
        lr = learning_rate.float().reshape(())
        for i, (u, p, p_copy) in enumerate(zip(*tensor_lists)):
            pn, un = per_tensor_param_norm[i], per_tensor_update_norm[update_norm_offset[i]]
            adaptive = (un != 0) & (pn != 0)
            if not use_nvlamb:
                adaptive = adaptive & (per_tensor_decay[i] != 0)
            ratio = torch.where(adaptive, lr * pn / torch.where(un != 0, un, torch.ones_like(un)), lr)
            pf = p.float() - ratio * u.float()
            p.copy_(pf)
            p_copy.copy_(pf)

    def lamb(chunk_size, noop_flag, tensor_lists, lr, beta1, beta2, epsilon, step, bias_correction, weight_decay, grad_averaging, mode,
             global_grad_norm, max_grad_norm):
        """fused_lamb_cuda.lamb: the deprecated contrib LAMB with a HOST float global_grad_norm (fused_lamb_cuda.cpp:3)."""
        dev = tensor_lists[0][0].device
        gn = torch.full((1,), float(global_grad_norm), dtype=torch.float32, device=dev)
        amp_C.multi_tensor_lamb(chunk_size, noop_flag, tensor_lists, lr, beta1, beta2, epsilon, step, bias_correction, weight_decay, grad_averaging,
                                mode, gn, max_grad_norm)

    return {
        "distributed_adam_cuda": _mod("distributed_adam_cuda", multi_tensor_fused_adam=multi_tensor_fused_adam,
                                      multi_tensor_fused_adam_capturable=multi_tensor_fused_adam_capturable,
                                      multi_tensor_fused_adam_with_param_remainders=multi_tensor_fused_adam_with_param_remainders),
        "distributed_lamb_cuda": _mod("distributed_lamb_cuda", multi_tensor_lamb_compute_update_term=multi_tensor_lamb_compute_update_term,
                                      multi_tensor_lamb_update_weights=multi_tensor_lamb_update_weights),
        "fused_lamb_cuda": _mod("fused_lamb_cuda", lamb=lamb),
    }



def _contrib_raw_mods():
    """Raw-extension names of the contrib packages whose entry points map one-to-one onto this library's kernels (the reference's Python
    modules drive exactly these calls): ``fused_conv_bias_relu`` (conv_bias_relu.py:13-94), ``group_norm_cuda`` / ``group_norm_v2_cuda``
    (group_norm.py:75-141), ``transducer_joint_cuda`` / ``transducer_loss_cuda`` (transducer.py:190-300), ``nccl_p2p_cuda``
    (csrc/nccl_p2p/nccl_p2p.cpp:20-28), ``_apex_nccl_allocator`` (NCCLAllocator.cpp:40), ``_apex_gpu_direct_storage`` (``_GDSFile``)."""
    mods = {}

    # ---- fused_conv_bias_relu: forward(inputs, padding, stride) -> [out]; backward([x, w, out, dy], ...) -> [dx, dw, db]
    from .contrib.conv_bias_relu import conv_bias_relu as CB

    def _conv_fwd(x, w, bias, scale, mask, padding, stride, relu):
        ctx = _Ctx()
        ctx.needs_input_grad = (False,) * 9
        with torch.no_grad():
            return CB.FusedConvEpilogue.forward(ctx, x, w, bias, scale, None, mask, stride, padding, relu)

    def _conv_bwd(x, w, out, dy, scale, padding, stride, relu, want_bias):
        dyc, xc, wc = CB._cl(dy), CB._cl(x), CB._cl(w)
        sc32 = scale.detach().reshape(-1).float().contiguous() if scale is not None else None
        g, _, dbias, _ = CB.epilogue_bwd(dyc, out if relu else None, None, None, sc32, relu, False, want_bias, False)
        dx, dw, _ = torch.ops.aten.convolution_backward(g, xc, wc.to(xc.dtype), None, CB._pair(stride), CB._pair(padding), (1, 1), False, (0, 0), 1,
                                                        (True, True, False))
        return [dx, dw.to(w.dtype)] + ([dbias.reshape(1, -1, 1, 1).to(w.dtype)] if want_bias else [])

    mods["fused_conv_bias_relu"] = _mod(
        "fused_conv_bias_relu",
        forward=lambda inputs, padding, stride: [_conv_fwd(inputs[0], inputs[1], inputs[2], None, None, padding, stride, True)],
        forward_mask=lambda inputs, padding, stride: [_conv_fwd(inputs[0], inputs[1], inputs[2], None, inputs[3], padding, stride, True)],
        forward_no_relu=lambda inputs, padding, stride: [_conv_fwd(inputs[0], inputs[1], inputs[2], None, None, padding, stride, False)],
        forward_cscale_cbias_relu=lambda inputs, padding, stride: [_conv_fwd(inputs[0], inputs[1], inputs[3], inputs[2], None, padding, stride, True)],
        backward=lambda inputs, padding, stride: _conv_bwd(inputs[0], inputs[1], inputs[2], inputs[3], None, padding, stride, True, True),
        backward_no_relu=lambda inputs, padding, stride: _conv_bwd(inputs[0], inputs[1], None, inputs[2], None, padding, stride, False, True),
        backward_cscale_cbias_relu=lambda inputs, padding, stride: _conv_bwd(inputs[0], inputs[1], inputs[3], inputs[4], inputs[2], padding, stride,
                                                                             True, False))

    # ---- group_norm_cuda (v1) / group_norm_v2_cuda: one kernel family serves both names
    from .contrib.group_norm import group_norm as GN

    def gn_forward(x, G, weight, bias, eps, passes, with_swish):
        y, sums = GN.group_norm_nhwc_fprop(x, G, weight, bias, eps, "silu" if with_swish else "")
        return y, sums.reshape(-1)

    def gn_backward(grad_output, sums, x, G, weight, bias, eps, passes, with_swish):
        return GN.group_norm_nhwc_bprop(grad_output, sums.reshape(2, -1), x, G, weight, bias, eps, "silu" if with_swish else "")

    def gn_v2(x, weight, bias, eps, with_swish, G, mean_var_out=None, sm_margin=0):
        y, sums = GN.group_norm_nhwc_fprop(x, G, weight, bias, eps, "silu" if with_swish else "")
        if mean_var_out is not None:
            mean_var_out.copy_(sums.reshape(-1))
        return y

    def gn_v2_bwd(grad_output, x, weight, bias, sums, eps, with_swish, G, sm_margin=0):
        return GN.group_norm_nhwc_bprop(grad_output, sums.reshape(2, -1), x, G, weight, bias, eps, "silu" if with_swish else "")

    mods["group_norm_cuda"] = _mod("group_norm_cuda", forward=gn_forward, backward=gn_backward)
    mods["group_norm_v2_cuda"] = _mod("group_norm_v2_cuda", gn=gn_v2, gn_bwd=gn_v2_bwd)

    # ---- transducer joint / loss
    from .contrib.transducer import transducer as TR

    def joint_forward(f, g, f_len, g_len, batch_offset, packed_batch, opt, pack_output, relu, dropout, dropout_prob, tile_size):
        ctx = _Ctx()
        p = float(dropout_prob) if dropout else 0.0
        with torch.no_grad():
            out = TR._JointFn.forward(ctx, f, g, f_len, g_len, batch_offset, packed_batch, bool(pack_output), bool(relu), p, None)
        mask = ctx.saved_tensors[0]
        return [out, mask.view(out.shape)] if mask is not None else [out]

    def joint_backward(inputs, f_len, g_len, batch_offset, max_f_len, max_g_len, pack_output, scale):
        dout = inputs[0]
        mask = inputs[1].reshape(-1).to(torch.uint8) if len(inputs) > 1 else None
        ctx = _Ctx()
        B, H = f_len.shape[0], dout.shape[-1]
        ctx.saved_tensors = (mask, TR._i32(f_len), TR._i32(g_len), batch_offset.to(torch.int64).contiguous() if pack_output else None)
        ctx.dims = (B, int(max_f_len), int(max_g_len), H, bool(pack_output), float(scale))
        with torch.no_grad():
            return list(TR._JointFn.backward(ctx, dout)[:2])

    def loss_forward(x, label, f_len, y_len, batch_offset, max_f_len, blank_idx, opt, packed_input):
        """x: log-probabilities (the reference applies log_softmax first): their log-sum-exp is 0, so the logits kernel computes the same
        alpha / beta / loss. -> (alpha, beta, loss)"""
        ctx = _Ctx()
        with torch.no_grad():
            loss = TR._LossFn.forward(ctx, x, label, f_len, y_len, batch_offset, max_f_len, blank_idx, bool(packed_input))
        return ctx.saved_tensors[2], ctx.saved_tensors[3], loss

    def loss_backward(x, loss_grad, alpha, beta, f_len, y_len, label, batch_offset, max_f_len, blank_idx, opt, fuse_softmax_backward, packed_input):
        if not fuse_softmax_backward:
            raise NotImplementedError("transducer_loss_cuda.backward: only the fused softmax backward (gradient w.r.t. the logits) is provided; "
                                      "use apex_b200.contrib.transducer.TransducerLoss for the unfused form")
        ctx = _Ctx()
        V = x.shape[-1]
        B, T, U = alpha.shape
        lse = torch.zeros(x.numel() // V, dtype=torch.float32, device=x.device)
        ctx.saved_tensors = (x.contiguous(), lse, alpha, beta, TR._i32(label), TR._i32(f_len), TR._i32(y_len),
                             batch_offset.to(torch.int64).contiguous() if packed_input else None)
        ctx.dims = (B, T, U, V, int(blank_idx), int(bool(packed_input)))
        with torch.no_grad():
            return TR._LossFn.backward(ctx, loss_grad)[0]

    mods["transducer_joint_cuda"] = _mod("transducer_joint_cuda", forward=joint_forward, backward=joint_backward)
    mods["transducer_loss_cuda"] = _mod("transducer_loss_cuda", forward=loss_forward, backward=loss_backward)

    # ---- communicator / allocator / storage modules: their Python modules already carry the raw entry points
    from .contrib.nccl_p2p import nccl_p2p as NP

    mods["nccl_p2p_cuda"] = _mod("nccl_p2p_cuda", get_unique_nccl_id=NP.get_unique_nccl_id, init_nccl_comm=NP.init_nccl_comm,
                                 left_right_halo_exchange_inplace=NP.left_right_halo_exchange_inplace,
                                 left_right_halo_exchange=NP.left_right_halo_exchange, add_delay=NP.add_delay)
    from .contrib import gpu_direct_storage as GDS

    class _GDSFile(GDS.GDSFile):
        """The reference's pybind class is opened by its constructor and closed by ``close()`` (gpu_direct_storage/__init__.py:17-21)."""

        def __init__(self, filename, mode):
            super().__init__(filename, mode)
            self.__enter__()

        def close(self):
            self.__exit__(None, None, None)

    mods["_apex_gpu_direct_storage"] = _mod("_apex_gpu_direct_storage", _GDSFile=_GDSFile)
    from .contrib.nccl_allocator import nccl_allocator as NA

    def get_nccl_allocator():
        """The pluggable allocator object behind ``create_nccl_mem_pool`` (reference NCCLAllocator.cpp:17-40): torch's NCCL ``mem_allocator``
        when the build has one; the symmetric-heap pool of this library otherwise."""
        import torch.distributed as dist

        try:
            backend = dist.distributed_c10d._get_default_group()._get_backend(torch.device("cuda"))
            return backend.mem_allocator
        except Exception:  # noqa: BLE001 - no process group / no NCCL allocator in this build
            return NA.create_symmetric_mem_pool()

    mods["_apex_nccl_allocator"] = _mod("_apex_nccl_allocator", get_nccl_allocator=get_nccl_allocator)
    mods["permutation_search_cuda"] = _perm_search_mod()
    mods["fmhalib"] = _fmhalib_mod()
    mods["cudnn_gbn_lib"] = _cudnn_gbn_mod()
    # ---- fast_multihead_attn: the 8 forward / backward pairs with the reference's intermediate-tensor conventions
    from .contrib.multihead_attn import raw_ext

    from .contrib.groupbn import raw_ext as bnp_ext

    mods["bnp"] = _mod("bnp", **{n: getattr(bnp_ext, n) for n in bnp_ext.ENTRY_POINTS})
    mods["peer_memory_cuda"] = _peer_memory_mod()
    mods["fast_bottleneck"] = _fast_bottleneck_mod()
    mods["fast_multihead_attn"] = _mod("fast_multihead_attn", **{n: getattr(raw_ext, n) for n in raw_ext.ENTRY_POINTS})
    return mods


def blob_strides(shape, channels_last: bool):
    """Strides of a typed view into a raw blob: dense row-major, or the NHWC strides of a logical [N, C, H, W] shape
    (reference peer_memory_cuda.cu:34-55)."""
    shape = [int(d) for d in shape]
    if channels_last:
        assert len(shape) == 4, "channels_last views are 4-D"
        n, c, h, w = shape
        return [c * h * w, 1, c * w, c]
    strides, acc = [], 1
    for d in reversed(shape):
        strides.append(acc)
        acc *= d
    return strides[::-1]


def _fast_bottleneck_mod():
    """``fast_bottleneck.forward / backward`` (reference apex/contrib/csrc/bottleneck/bottleneck.cpp:1004-1141, 1143-1378): the whole
    ResNet bottleneck on explicit tensor lists, as ``BottleneckFunction`` of the reference drives them (bottleneck.py:80-132).
    forward(nhwc, stride, [x, w1, w2, w3, s1, s2, s3, b1, b2, b3(, w4, s4, b4)]) -> [out1, out2, out3];
    backward(nhwc, stride, [x, w1, w2, w3, s1, s2, s3, b1, b2, b3, grad_conv3, grad_conv4, out1, out2(, w4)]) -> [dx, dw1, dw2, dw3(, dw4)]
    where grad_conv3 / grad_conv4 are the caller's drelu x scale products of the main and the identity branch. Library convolutions
    (as in the reference) around the fused scale-bias-(add)-ReLU tail of contrib/conv_bias_relu. The 21 staged entry points of the
    spatial-parallel pipeline (forward_out2_halo, backward_grad_out1_halo_corr, ...) have no counterpart: ``SpatialBottleneck`` here
    runs one halo-overlapped convolution Function instead of a hand-staged cuDNN graph sequence."""
    from torch.nn import grad as G

    from .contrib.conv_bias_relu.conv_bias_relu import fused_conv_epilogue

    def _nchw(t, nhwc):
        return t.permute(0, 3, 1, 2) if nhwc else t

    def _back(t, nhwc):
        return t.permute(0, 2, 3, 1).contiguous() if nhwc else t

    vec = lambda t: t.reshape(1, -1, 1, 1)                                              # noqa: E731

    def forward(explicit_nhwc, stride_1X1, inputs):
        x, w1, w2, w3, s1, s2, s3, b1, b2, b3 = inputs[:10]
        x, w1, w2, w3 = (_nchw(t, explicit_nhwc) for t in (x, w1, w2, w3))
        with torch.no_grad():
            out1 = fused_conv_epilogue(x, w1, bias=vec(b1), scale=vec(s1), stride=stride_1X1, padding=0, relu=True)
            out2 = fused_conv_epilogue(out1, w2, bias=vec(b2), scale=vec(s2), stride=1, padding=1, relu=True)
            identity = x
            if len(inputs) > 10:
                w4, s4, b4 = inputs[10:13]
                identity = fused_conv_epilogue(x, _nchw(w4, explicit_nhwc), bias=vec(b4), scale=vec(s4), stride=stride_1X1, padding=0,
                                               relu=False)
            out3 = fused_conv_epilogue(out2, w3, bias=vec(b3), scale=vec(s3), z=identity, stride=1, padding=0, relu=True)
        return [_back(o, explicit_nhwc) for o in (out1, out2, out3)]

    def backward(explicit_nhwc, stride_1X1, inputs):
        x, w1, w2, w3, s1, s2, s3 = inputs[:7]
        g3, g4, out1, out2 = inputs[10:14]
        x, w1, w2, w3, g3, g4, out1, out2 = (_nchw(t, explicit_nhwc) for t in (x, w1, w2, w3, g3, g4, out1, out2))
        dw3 = G.conv2d_weight(out2, w3.shape, g3)
        g2 = G.conv2d_input(out2.shape, w3, g3) * (out2 > 0).to(g3.dtype) * vec(s2).to(g3.dtype)
        dw2 = G.conv2d_weight(out1, w2.shape, g2, padding=1)
        g1 = G.conv2d_input(out1.shape, w2, g2, padding=1) * (out1 > 0).to(g3.dtype) * vec(s1).to(g3.dtype)
        dw1 = G.conv2d_weight(x, w1.shape, g1, stride=stride_1X1)
        dx = G.conv2d_input(x.shape, w1, g1, stride=stride_1X1)
        grads = [dw1, dw2, dw3]
        if len(inputs) > 14:
            w4 = _nchw(inputs[14], explicit_nhwc)
            dx = dx + G.conv2d_input(x.shape, w4, g4, stride=stride_1X1)
            grads.append(G.conv2d_weight(x, w4.shape, g4, stride=stride_1X1))
        else:
            dx = dx + g4
        return [_back(t, explicit_nhwc) for t in [dx] + grads]

    return _mod("fast_bottleneck", forward=forward, backward=backward)


def _peer_memory_mod():
    """``peer_memory_cuda`` (reference apex/contrib/csrc/peer_memory/peer_memory.cpp:19-36): raw cudaIpc blobs addressed by integer
    pointers. ``PeerMemoryPool`` here lives on the VMM symmetric heap and needs none of this; the raw calls serve code written against
    the extension: blobs come from the cudaIpc functions of csrc/symm_heap.cpp, the halo exchange is staged as copy -> barrier -> copy
    (the one-kernel exchange, csrc/halo_exchange.cu, needs the heap's signal pads, which a bare pointer API cannot name)."""
    import ctypes

    from . import _lib
    from .parallel.symmetric import _tensor_from_ptr

    _lib.declare("ab_ipc_alloc", "i l p p")
    _lib.declare("ab_ipc_open", "i p p")
    _lib.declare("ab_ipc_free", "l")
    blobs: dict = {}                                                   # raw pointer -> (ipc handle bytes, nbytes)

    def _need_gpu(what):
        if not (_lib.available() and torch.cuda.is_available()):
            raise RuntimeError(f"peer_memory_cuda.{what} needs a CUDA device and the apex_b200 native library")

    def _bytes_view(raw, nbytes):
        return _tensor_from_ptr(int(raw), int(nbytes), torch.device("cuda", torch.cuda.current_device()), None)

    def allocate_raw(size):
        _need_gpu("allocate_raw")
        p, h = ctypes.c_uint64(0), ctypes.create_string_buffer(64)
        _lib.fn("ab_ipc_alloc")(torch.cuda.current_device(), int(size), ctypes.addressof(p), ctypes.addressof(h))
        blobs[int(p.value)] = (bytes(h.raw), int(size))
        _bytes_view(p.value, size).zero_()
        return int(p.value)

    def free_raw(raw):
        _need_gpu("free_raw")
        torch.cuda.synchronize()
        blobs.pop(int(raw), None)
        _lib.fn("ab_ipc_free")(int(raw))

    def zero(raw, size):
        _need_gpu("zero")
        _bytes_view(raw, size).zero_()

    def get_raw_ipc_address(raw):
        if int(raw) not in blobs:
            raise ValueError("get_raw_ipc_address: not a pointer returned by allocate_raw")
        return torch.frombuffer(bytearray(blobs[int(raw)][0]), dtype=torch.uint8).clone()

    def get_raw_peers(ipc_addresses, peer_rank, raw):
        _need_gpu("get_raw_peers")
        out = []
        for i in range(ipc_addresses.size(0)):
            if i == int(peer_rank):
                out.append(int(raw))
                continue
            h = ctypes.create_string_buffer(bytes(ipc_addresses[i].cpu().contiguous().numpy().tobytes()), 64)
            p = ctypes.c_uint64(0)
            _lib.fn("ab_ipc_open")(torch.cuda.current_device(), ctypes.addressof(h), ctypes.addressof(p))
            out.append(int(p.value))
        return out

    def _blob_view(dtype):
        def view(raw, shape, channels_last):
            _need_gpu("blob_view")
            numel = 1
            for d in shape:
                numel *= int(d)
            flat = _bytes_view(raw, numel * torch.empty((), dtype=dtype).element_size()).view(dtype)
            return flat.as_strided([int(d) for d in shape], blob_strides(shape, channels_last))
        return view

    def push_pull_halos_1d(diagnostics, explicit_nhwc, numSM, rank, top_zero, top_out_halo, top_in_transfer, top_out_transfer,
                           top_in_halo, btm_zero, btm_out_halo, btm_in_transfer, btm_out_transfer, btm_in_halo):
        """Send the two outgoing halos into the neighbours' transfer buffers, receive the incoming ones from the local transfer buffers
        (or zero them at the ends of the split). Collective over the default process group."""
        import torch.distributed as dist

        assert not (top_zero and btm_zero)
        if not top_zero:
            top_out_transfer.copy_(top_out_halo)
        if not btm_zero:
            btm_out_transfer.copy_(btm_out_halo)
        torch.cuda.current_stream().synchronize()
        if dist.is_initialized():
            dist.barrier()                                   # every push has landed
        top_in_halo.zero_() if top_zero else top_in_halo.copy_(top_in_transfer)
        btm_in_halo.zero_() if btm_zero else btm_in_halo.copy_(btm_in_transfer)
        torch.cuda.current_stream().synchronize()
        if dist.is_initialized():
            dist.barrier()                                   # transfer buffers may be overwritten by the next exchange

    return _mod("peer_memory_cuda", allocate_raw=allocate_raw, free_raw=free_raw, zero=zero, get_raw_ipc_address=get_raw_ipc_address,
                get_raw_peers=get_raw_peers, blob_view_half=_blob_view(torch.float16), blob_view_float=_blob_view(torch.float32),
                blob_view_int=_blob_view(torch.int32), push_pull_halos_1d=push_pull_halos_1d)


_gbn_groups: dict = {}


def _cudnn_gbn_mod():
    """``cudnn_gbn_lib`` (reference apex/contrib/cudnn_gbn/batch_norm.py:34-69 over apex/contrib/csrc/cudnn_gbn): NHWC batch norm whose
    statistics span ``group_size`` consecutive ranks. ``forward`` fills ``minibatch_mean`` / ``minibatch_inv_var`` and updates the running
    statistics in place; ``backward`` returns ``(dx, dscale, dbias)``. Composed from the SyncBN kernel's phase ops
    (``apex_b200.parallel.syncbn_ops``: Welford statistics, normalise, backward reductions) plus one all-gather / all-reduce of the
    per-channel vectors over the group's process group; the reference's ``peer_buffers`` lists (its hand-rolled exchange areas) are
    accepted and not used."""
    import torch.distributed as dist

    from .parallel import syncbn_ops as S

    def _group(group_size, group_rank):
        if group_size <= 1 or not (dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(), dist.get_rank()
        if group_size >= world:
            return dist.group.WORLD
        key = (world, group_size)
        if key not in _gbn_groups:      # every rank creates every group, in the same order (new_group is collective)
            _gbn_groups[key] = [dist.new_group(list(range(g0, g0 + group_size))) for g0 in range(0, world, group_size)]
        return _gbn_groups[key][rank // group_size]

    def forward(x, weight, bias, running_mean, running_var, minibatch_mean, minibatch_inv_var, momentum, eps, group_size, group_rank, peer_buffers):
        mean, var_b = S.welford_mean_var(x)
        count = x.numel() // x.shape[1]
        pg = _group(int(group_size), int(group_rank))
        if pg is not None:
            n = dist.get_world_size(pg)
            packed = torch.cat([mean, var_b, torch.full((1,), float(count), device=mean.device)])
            gathered = [torch.empty_like(packed) for _ in range(n)]
            dist.all_gather(gathered, packed, group=pg)
            allm = torch.stack(gathered)
            C = mean.numel()
            mean, var_u, inv_std = S.welford_parallel(allm[:, :C], allm[:, C:2 * C], allm[:, 2 * C], float(eps))
        else:
            var_u = var_b * (count / max(count - 1, 1))
            inv_std = torch.rsqrt(var_b + float(eps))
        with torch.no_grad():
            minibatch_mean.copy_(mean.reshape(minibatch_mean.shape))
            minibatch_inv_var.copy_(inv_std.reshape(minibatch_inv_var.shape))
            if running_mean is not None:
                running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                running_var.mul_(1 - momentum).add_(var_u.to(running_var.dtype), alpha=momentum)
        return S.batchnorm_forward(x, mean, inv_std, weight, bias)

    def backward(x, dy, scale, minibatch_mean, minibatch_inv_var, eps, group_size, group_rank, peer_buffers):
        mean, inv_std = minibatch_mean.reshape(-1).float(), minibatch_inv_var.reshape(-1).float()
        sum_dy, sum_dy_xmu, dscale, dbias = S.reduce_bn(dy, x, mean, inv_std, scale)
        count = torch.tensor([float(x.numel() // x.shape[1])], device=x.device)
        pg = _group(int(group_size), int(group_rank))
        if pg is not None:
            packed = torch.cat([sum_dy, sum_dy_xmu, count])
            dist.all_reduce(packed, group=pg)
            C = sum_dy.numel()
            sum_dy, sum_dy_xmu, count = packed[:C], packed[C:2 * C], packed[2 * C:]
        dx = S.batchnorm_backward(dy, x, mean, inv_std, scale, sum_dy, sum_dy_xmu, count)
        return dx, dscale, dbias

    return _mod("cudnn_gbn_lib", forward=forward, backward=backward)


def _fmhalib_mod():
    """``fmhalib`` (reference apex/contrib/csrc/fmha/fmha_api.cpp:86-127 as driven by apex/contrib/fmha/fmha.py:47-90): ``fwd`` / ``fwd_nl``
    return ``(context, S_dmask)`` and ``bwd`` / ``bwd_nl`` take ``S_dmask`` back. The reference materialises the dropout-masked softmax in
    ``S_dmask``; the kernels here recompute the probabilities from the log-sum-exp, so ``S_dmask`` is an OPAQUE fp32 state tensor (row
    log-sum-exp, the forward output and the Philox counters) that is only meaningful when handed back to ``bwd``. No sequence-length or
    head-dimension limit of the reference applies (head dim 64 / 128 on the tcgen05 kernels)."""
    from .contrib.fmha import kernels as K

    def _pack(lse, out, philox):
        meta = torch.tensor([philox[0] & 0xFFFFFFFF, (philox[0] >> 32) & 0xFFFFFFFF, philox[1] & 0xFFFFFFFF, (philox[1] >> 32) & 0xFFFFFFFF,
                             lse.numel(), out.numel()], dtype=torch.int64, device=lse.device).to(torch.int32)
        return torch.cat([meta.view(torch.float32), lse.reshape(-1).float(), out.reshape(-1).float()])

    def _unpack(state, like):
        meta = state[:6].view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        m = [int(v) for v in meta.tolist()]
        n_lse, n_out = m[4], m[5]
        total, h, d = like.shape[0], like.shape[2], like.shape[3]
        lse = state[6:6 + n_lse].view(total, h)
        out = state[6 + n_lse:6 + n_lse + n_out].view(total, h, d).to(like.dtype)
        return lse, out, (m[0] | (m[1] << 32), m[2] | (m[3] << 32))

    def fwd(qkv, cu_seqlens, p_dropout, max_s, is_training, is_nl, zero_tensors, generator=None):
        total, three, h, d = qkv.shape
        if not K.supported(qkv, d):
            raise RuntimeError("fmhalib: fp16 / bf16 CUDA tensors with head dim 64 or 128 are required (use apex_b200.contrib.fmha.FMHA for the generic path)")
        p = float(p_dropout) if is_training else 0.0
        cu = cu_seqlens if cu_seqlens.dtype == torch.int32 else cu_seqlens.to(torch.int32)
        philox = K.next_philox(qkv.device) if p > 0.0 else (0, 0)
        out, lse = K.fmha_fwd(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu_seqlens_q=cu, cu_seqlens_k=cu, max_seqlen_q=int(max_s), seqlen_k=int(max_s),
                              return_lse=True, dropout_p=p, philox=philox)
        return out, _pack(lse, out, philox)

    def bwd(dout, qkv, S_dmask, cu_seqlens, p_dropout, max_s, zero_tensors):
        lse, out, philox = _unpack(S_dmask, qkv)
        cu = cu_seqlens if cu_seqlens.dtype == torch.int32 else cu_seqlens.to(torch.int32)
        p = float(p_dropout) if philox != (0, 0) else 0.0
        dq, dk, dv = K.fmha_bwd(dout.contiguous(), qkv[:, 0], qkv[:, 1], qkv[:, 2], out, lse, cu_seqlens_q=cu, cu_seqlens_k=cu, max_seqlen_q=int(max_s),
                                max_seqlen_k=int(max_s), dropout_p=p, philox=philox)
        return torch.stack([dq, dk, dv], 1), None

    def fwd_nl(qkv, cu_seqlens, p_dropout, max_s, is_training, is_nl, zero_tensors, generator=None):
        return fwd(qkv, cu_seqlens, p_dropout, max_s, is_training, is_nl, zero_tensors, generator)

    def bwd_nl(dout, qkv, S_dmask, cu_seqlens, p_dropout, max_s, zero_tensors):
        dqkv, dp = bwd(dout, qkv, S_dmask, cu_seqlens, p_dropout, max_s, zero_tensors)
        return dqkv, dp, None

    return _mod("fmhalib", fwd=fwd, bwd=bwd, fwd_nl=fwd_nl, bwd_nl=bwd_nl)


def _perm_search_mod():
    """``permutation_search_cuda`` (reference apex/contrib/sparsity/permutation_search_kernels/CUDA_kernels/permutation_search_kernels.cu:
    625-631): four entry points over flat numpy buffers; results are written into the caller's output arrays and 0 is returned. The launch
    geometry arguments of ``sum_after_2_to_4`` (blocks, threads) are accepted and ignored. Runs on csrc/perm_search.cu when a GPU is
    present, on the PyTorch forms of the same scores otherwise."""
    import numpy as np

    from .contrib.sparsity import permutation_search as PS

    def _dev():
        from . import _lib

        return torch.device("cuda") if (torch.cuda.is_available() and _lib.available()) else torch.device("cpu")

    def _matrix(buf, rows, cols):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(buf, dtype=np.float32))).view(int(rows), int(cols)).to(_dev())

    def sum_after_2_to_4(matrix, rows, cols, start_col, end_col, blocks, threads, output):
        m = _matrix(matrix, rows, cols)[:, int(start_col):int(end_col)].contiguous()
        output[0] = float(PS.sum_after_2_to_4(m))
        return 0

    def build_permute_map(matrix, rows, cols, stripes, num_groups, group_width, permutations, perm_length, improvements, best_indices):
        m = _matrix(matrix, rows, cols)
        groups = torch.from_numpy(np.asarray(stripes, dtype=np.int64).reshape(int(num_groups), int(group_width)).astype(np.int32)).to(m.device)
        cands = np.asarray(permutations, dtype=np.int64).reshape(-1, int(perm_length)).astype(np.uint8 if int(perm_length) <= 255 else np.int32)
        imp, idx = PS._score_groups(m, groups, cands)
        improvements[:int(num_groups)] = imp.float().cpu().numpy()
        best_indices[:int(num_groups)] = idx.cpu().numpy().astype(np.uint32)
        return 0

    def check_permutations(matrix, rows, cols, stripe_groups, group_width, num_groups, permutations, num_permutations, improvement, permutation):
        m = _matrix(matrix, rows, cols)
        sg = np.asarray(stripe_groups, dtype=np.int64).reshape(int(num_groups), int(group_width))
        perms = np.asarray(permutations, dtype=np.int64).reshape(int(num_permutations), -1)
        for g in range(int(num_groups)):
            colsel = torch.from_numpy((sg[g][:, None] * 4 + np.arange(4)).reshape(-1)).to(m.device)
            sub = m[:, colsel].contiguous()
            base = float(PS.sum_after_2_to_4(sub))
            scores = PS.sum_after_2_to_4(sub, torch.from_numpy(perms.astype(np.int32)).to(m.device))
            best = int(torch.argmax(scores))
            improvement[g] = float(scores[best]) - base
            permutation[g] = best
        return 0

    def build_swap_map(matrix, rows, cols, stripe_pairs, output):
        m = _matrix(matrix, rows, cols).abs()
        pairs = np.asarray(stripe_pairs, dtype=np.int64).reshape(-1, 2)
        kept = lambda t: t.reshape(t.shape[0], -1, 4).topk(2, dim=-1).values.sum()   # noqa: E731
        for i, (s0, s1) in enumerate(pairs):
            c = torch.cat([m[:, s0 * 4:s0 * 4 + 4], m[:, s1 * 4:s1 * 4 + 4]], 1)    # [rows, 8]
            base = float(kept(c))
            for k in range(16):
                a, b = k // 4, 4 + k % 4
                sw = c.clone()
                sw[:, [a, b]] = c[:, [b, a]]
                output[i * 16 + k] = float(kept(sw)) - base
        return 0

    return _mod("permutation_search_cuda", sum_after_2_to_4=sum_after_2_to_4, build_permute_map=build_permute_map,
                check_permutations=check_permutations, build_swap_map=build_swap_map)


class _Ctx:
    """Stand-in for an autograd context when a Function's forward / backward is driven directly."""

    needs_input_grad = (True,) * 8

    def save_for_backward(self, *t):
        self.saved_tensors = t


def extension_modules() -> dict:
    from .ops import amp_C
    from .parallel import syncbn_ops
    from .transformer.functional import fused_weight_gradient
    from .utils import flatten

    mods = {"amp_C": amp_C, "syncbn": syncbn_ops, "fused_weight_gradient_mlp_cuda": fused_weight_gradient,
            "apex_C": _mod("apex_C", flatten=flatten.flatten, unflatten=flatten.unflatten),
            "fused_layer_norm_cuda": _layer_norm_mod(), "fused_rotary_positional_embedding": _rope_mod(), "xentropy_cuda": _xentropy_mod()}
    mods.update(_softmax_mods())
    mods["fast_layer_norm"] = _fast_layer_norm_mod(mods["fused_layer_norm_cuda"])
    for group in (_dense_mods, _small_contrib_mods, _distopt_mods):
        mods.update(group())
    try:
        mods.update(_contrib_raw_mods())
    except Exception as e:  # noqa: BLE001 - a contrib package that cannot be imported here must not take the core names down
        import warnings

        warnings.warn(f"apex_b200.ext_compat: contrib extension names not registered ({type(e).__name__}: {e})")
    try:
        from .contrib.optimizers import fused_adam_cuda

        mods["fused_adam_cuda"] = fused_adam_cuda
    except Exception:  # noqa: BLE001
        pass
    return mods


def install(overwrite: bool = False) -> list[str]:
    """Register the extension-name modules in ``sys.modules`` (existing entries are kept unless ``overwrite``); returns the names."""
    done = []
    for name, m in extension_modules().items():
        if overwrite or name not in sys.modules:
            sys.modules[name] = m
            done.append(name)
    return done
