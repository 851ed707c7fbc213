"""RNN-T joint and loss. Reference: apex/contrib/transducer/transducer.py:6-195 over transducer_joint_cuda / transducer_loss_cuda
(apex/contrib/csrc/transducer/*.cu). Same module signatures and semantics (padding mask from f_len / g_len, optional packed output
with ``batch_offset``, ReLU / dropout on the joint, blank index, per-utterance negative log-likelihood).

The joint is a broadcast add (+ReLU, +dropout) in one fused elementwise expression; the loss runs the alpha recursion over
anti-diagonals of the (T, U) lattice, vectorised over the batch and the diagonal, in log space — T+U-1 small steps instead of a
thread-per-cell CUDA kernel. Gradients come from autograd through the same recursion (exactly the beta recursion)."""
from __future__ import annotations

import torch


class TransducerJoint(torch.nn.Module):
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


class TransducerLoss(torch.nn.Module):
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
