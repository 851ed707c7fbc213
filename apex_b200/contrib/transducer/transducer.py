"""RNN-T joint and loss. Reference: apex/contrib/transducer/transducer.py:6-195 over transducer_joint_cuda / transducer_loss_cuda
(apex/contrib/csrc/transducer/*.cu). Same module signatures and semantics (padding mask from f_len / g_len, optional packed output
with ``batch_offset``, ReLU / dropout on the joint, blank index, per-utterance negative log-likelihood).

On CUDA tensors both modules run the sm_100a kernels of csrc/transducer.cu (joint: one vectorised pass + one launch for both
backward reductions; loss: log-sum-exp per lattice cell, alpha/beta along anti-diagonals with a CTA per (utterance, direction),
backward fused with the softmax backward, no materialised log-softmax). On CPU tensors the ``_Torch*`` oracles below run: the same
math in PyTorch ops (anti-diagonal alpha recursion vectorised over the batch; gradients through autograd)."""
from __future__ import annotations

import torch

from ... import _lib

_lib.declare("ab_transducer_joint_fwd", "p p p p p p p i i i i i l i f l i p")
_lib.declare("ab_transducer_joint_bwd", "p p p p p p p i i i i i f i p")
_lib.declare("ab_transducer_loss_fwd", "p p p p p p p p p i i i i i i l i p")
_lib.declare("ab_transducer_loss_bwd", "p p p p p p p p p p i i i i i i i p")


class _TorchTransducerJoint(torch.nn.Module):
    def __init__(self, pack_output=False, relu=False, dropout=False, opt=1, fwd_tile_size=4, dropout_prob=0, probe_mask=False):
        super().__init__()
        self.pack_output, self.relu, self.dropout, self.dropout_prob = pack_output, relu, dropout, dropout_prob
        self.probe_mask = probe_mask
        self.mask_probe = []

    def forward(self, f, g, f_len, g_len, batch_offset=None, packed_batch=0):
        """f [B, T, H], g [B, U, H] -> [B, T, U, H], or packed [sum_b f_len[b]*g_len[b], H] when pack_output."""
        h = f.unsqueeze(2) + g.unsqueeze(1)
        if self.relu:
            h = torch.relu(h)
        if self.dropout and self.training and self.dropout_prob > 0:
            mask = torch.rand_like(h, dtype=torch.float32) >= self.dropout_prob
            if self.probe_mask:
                self.mask_probe = [mask]
            h = h * mask.to(h.dtype) / (1.0 - self.dropout_prob)
        B, T, U, _ = h.shape
        t_ok = torch.arange(T, device=f.device).view(1, T, 1) < f_len.view(B, 1, 1)
        u_ok = torch.arange(U, device=f.device).view(1, 1, U) < g_len.view(B, 1, 1)
        valid = t_ok & u_ok
        if self.pack_output:
            if batch_offset is None or packed_batch == 0:
                raise Exception("Please specify batch_offset and packed_batch when packing is enabled")
            return h[valid]
        return h * valid.unsqueeze(-1).to(h.dtype)


class _TorchTransducerLoss(torch.nn.Module):
    def __init__(self, fuse_softmax_backward=True, opt=1, packed_input=False):
        super().__init__()
        self.packed_input = packed_input

    def forward(self, x, label, f_len, y_len, blank_idx, batch_offset=None, max_f_len=None, debug_list=None):
        """x: joint logits [B, T, U, V] (or packed [N, V] with batch_offset / max_f_len); label [B, U-1]; returns -log p(y|x) per utterance."""
        if self.packed_input:
            if batch_offset is None or max_f_len is None:
                raise Exception("Please specify batch_offset and max_f_len when packing is enabled")
            B, U = label.shape[0], label.shape[1] + 1
            T = int(max_f_len)
            V = x.shape[-1]
            dense = x.new_zeros(B, T, U, V)
            t_ok = torch.arange(T, device=x.device).view(1, T, 1) < f_len.view(B, 1, 1)
            u_ok = torch.arange(U, device=x.device).view(1, 1, U) < (y_len + 1).view(B, 1, 1)
            dense[t_ok & u_ok] = x
            x = dense
        B, T, U, V = x.shape
        lp = torch.log_softmax(x.float(), dim=-1)
        blank = lp[..., blank_idx]                                             # [B, T, U]
        lab = label.long().clamp(min=0)
        emit = lp[:, :, :U - 1, :].gather(3, lab.view(B, 1, U - 1, 1).expand(B, T, U - 1, 1)).squeeze(3)  # [B, T, U-1]
        neg = torch.finfo(torch.float32).min / 4
        alpha = torch.full((B, T, U), neg, device=x.device, dtype=torch.float32)
        alpha[:, 0, 0] = 0.0
        rows = []
        # anti-diagonal sweep: cells with t + u = d depend only on diagonal d-1
        cur = alpha[:, 0:1, 0]  # placeholder; we keep the full lattice as a list of diagonals for autograd friendliness
        diag = {0: alpha[:, 0, 0].unsqueeze(1)}  # d -> [B, n_cells] ordered by t
        for d in range(1, T + U - 1):
            t_lo, t_hi = max(0, d - (U - 1)), min(T - 1, d)
            ts = torch.arange(t_lo, t_hi + 1, device=x.device)
            us = d - ts
            prev = diag[d - 1]
            p_lo = max(0, d - 1 - (U - 1))
            # from (t-1, u): needs t >= 1
            a_t = torch.full((B, ts.numel()), neg, device=x.device)
            has_t = ts >= 1
            if has_t.any():
                tt, uu = ts[has_t] - 1, us[has_t]
                keep = uu <= U - 1
                a_t[:, has_t] = prev[:, (tt - p_lo)] + blank[:, tt, uu]
            a_u = torch.full((B, ts.numel()), neg, device=x.device)
            has_u = us >= 1
            if has_u.any():
                tt, uu = ts[has_u], us[has_u] - 1
                a_u[:, has_u] = prev[:, (tt - p_lo)] + emit[:, tt, uu]
            diag[d] = torch.logaddexp(a_t, a_u)
        # log-likelihood: alpha[f_len-1, y_len] + blank[f_len-1, y_len]
        out = []
        for b in range(B):
            t, u = int(f_len[b]) - 1, int(y_len[b])
            d = t + u
            t_lo = max(0, d - (U - 1))
            out.append(-(diag[d][b, t - t_lo] + blank[b, t, u]))
        return torch.stack(out).to(x.dtype if x.dtype == torch.float32 else torch.float32)


def _i32(t):
    return t.to(torch.int32).contiguous()


class _JointFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, g, f_len, g_len, batch_offset, packed_batch, pack, relu, drop_p, probe):
        f, g = f.contiguous(), g.contiguous()
        B, T, H = f.shape
        U = g.shape[1]
        fl, gl = _i32(f_len), _i32(g_len)
        bo = batch_offset.to(torch.int64).contiguous() if pack else None
        rows = int(packed_batch) if pack else B * T * U
        out = torch.empty((rows, H) if pack else (B, T, U, H), dtype=f.dtype, device=f.device)
        need_mask = relu or drop_p > 0
        mask = torch.empty(rows * H, dtype=torch.uint8, device=f.device) if need_mask else None
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0 else 0
        _lib.fn("ab_transducer_joint_fwd")(f.data_ptr(), g.data_ptr(), out.data_ptr(), _lib.ptr(mask), fl.data_ptr(), gl.data_ptr(),
                                           _lib.ptr(bo), B, T, U, H, int(pack), rows, int(relu), float(drop_p), seed, _lib.dt(f),
                                           _lib.stream_ptr(f.device))
        ctx.save_for_backward(mask, fl, gl, bo)
        ctx.dims = (B, T, U, H, pack, 1.0 / (1.0 - drop_p) if drop_p > 0 else 1.0)
        if probe is not None and mask is not None:
            probe.append(mask.view(out.shape).bool())
        return out

    @staticmethod
    def backward(ctx, dout):
        mask, fl, gl, bo = ctx.saved_tensors
        B, T, U, H, pack, scale = ctx.dims
        dout = dout.contiguous()
        df = torch.empty(B, T, H, dtype=dout.dtype, device=dout.device)
        dg = torch.empty(B, U, H, dtype=dout.dtype, device=dout.device)
        _lib.fn("ab_transducer_joint_bwd")(dout.data_ptr(), _lib.ptr(mask), df.data_ptr(), dg.data_ptr(), fl.data_ptr(), gl.data_ptr(),
                                           _lib.ptr(bo), B, T, U, H, int(pack), float(scale), _lib.dt(dout), _lib.stream_ptr(dout.device))
        return df, dg, None, None, None, None, None, None, None, None


class TransducerJoint(_TorchTransducerJoint):
    """f [B, T, H] + g [B, U, H] -> [B, T, U, H] (padding zeroed) or packed [sum_b f_len*g_len, H] (reference transducer.py:6-86)."""

    def forward(self, f, g, f_len, g_len, batch_offset=None, packed_batch=0):
        if not (f.is_cuda and _lib.available()):
            return super().forward(f, g, f_len, g_len, batch_offset, packed_batch)
        if self.pack_output and (batch_offset is None or packed_batch == 0):
            raise Exception("Please specify batch_offset and packed_batch when packing is enabled")
        p = float(self.dropout_prob) if (self.dropout and self.training) else 0.0
        self.mask_probe = []
        return _JointFn.apply(f, g, f_len, g_len, batch_offset, packed_batch, self.pack_output, self.relu, p,
                              self.mask_probe if self.probe_mask else None)


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, label, f_len, y_len, batch_offset, max_f_len, blank_idx, packed):
        x = x.contiguous()
        B, U = label.shape[0], label.shape[1] + 1
        T = int(max_f_len) if packed else x.shape[1]
        V = x.shape[-1]
        rows = x.numel() // V
        dev = x.device
        lab, fl, yl = _i32(label), _i32(f_len), _i32(y_len)
        bo = batch_offset.to(torch.int64).contiguous() if packed else None
        lse = torch.empty(rows, dtype=torch.float32, device=dev)
        alpha = torch.empty(B, T, U, dtype=torch.float32, device=dev)
        beta = torch.empty(B, T, U, dtype=torch.float32, device=dev)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        _lib.fn("ab_transducer_loss_fwd")(x.data_ptr(), lse.data_ptr(), alpha.data_ptr(), beta.data_ptr(), loss.data_ptr(), lab.data_ptr(),
                                          fl.data_ptr(), yl.data_ptr(), _lib.ptr(bo), B, T, U, V, int(blank_idx), int(packed), rows,
                                          _lib.dt(x), _lib.stream_ptr(dev))
        ctx.save_for_backward(x, lse, alpha, beta, lab, fl, yl, bo)
        ctx.dims = (B, T, U, V, int(blank_idx), int(packed))
        return loss

    @staticmethod
    def backward(ctx, gloss):
        x, lse, alpha, beta, lab, fl, yl, bo = ctx.saved_tensors
        B, T, U, V, blank, packed = ctx.dims
        dx = torch.empty_like(x)
        g = gloss.float().contiguous()
        _lib.fn("ab_transducer_loss_bwd")(x.data_ptr(), lse.data_ptr(), alpha.data_ptr(), beta.data_ptr(), g.data_ptr(), dx.data_ptr(),
                                          lab.data_ptr(), fl.data_ptr(), yl.data_ptr(), _lib.ptr(bo), B, T, U, V, blank, packed,
                                          _lib.dt(x), _lib.stream_ptr(x.device))
        return dx, None, None, None, None, None, None, None


class TransducerLoss(_TorchTransducerLoss):
    """-log p(y | x) per utterance from joint logits x [B, T, U, V] (or packed [N, V]) (reference transducer.py:88-195)."""

    def forward(self, x, label, f_len, y_len, blank_idx, batch_offset=None, max_f_len=None, debug_list=None):
        if not (x.is_cuda and _lib.available()):
            return super().forward(x, label, f_len, y_len, blank_idx, batch_offset, max_f_len, debug_list)
        if self.packed_input and (batch_offset is None or max_f_len is None):
            raise Exception("Please specify batch_offset and max_f_len when packing is enabled")
        return _LossFn.apply(x, label, f_len, y_len, batch_offset, max_f_len, blank_idx, self.packed_input)


# reference class names of the autograd functions (transducer.py:197-436)
TransducerJointFunc = _JointFn
TransducerLossFunc = _LossFn
