"""MLP(mlp_sizes, bias, activation in {none, relu, sigmoid}) — N linear layers with the activation fused into the GEMM epilogue.

Reference: apex/mlp/mlp.py:33-87 over csrc/mlp_cuda.cu (cuBLAS GEMM + separate bias/activation kernels unless activation<1).
Here every layer is one tcgen05 GEMM whose epilogue applies bias + ReLU/sigmoid; backward recomputes the activation
derivative from the saved layer OUTPUT (relu: y>0, sigmoid: y(1-y)), so only layer outputs are stashed.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from ..ops import gemm as G

_ACT = {"none": 0, "relu": 1, "sigmoid": 2}


class MlpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bias, activation, input, *params):
        n = len(params) // (2 if bias else 1)
        weights, biases = params[:n], (params[n:] if bias else [None] * n)
        x = input.reshape(-1, input.shape[-1]).contiguous()
        acts = [x]
        for w, b in zip(weights, biases):
            if activation == 1:
                epi = G.EPI_BIAS_RELU if b is not None else G.EPI_RELU
            elif activation == 2:
                epi = G.EPI_BIAS_SIGMOID if b is not None else G.EPI_SIGMOID
            else:
                epi = G.EPI_BIAS if b is not None else G.EPI_NONE
            acts.append(G.linear_fwd(acts[-1], w.contiguous(), b, epi=epi))
        ctx.save_for_backward(*acts, *weights)
        ctx.n, ctx.bias, ctx.activation, ctx.in_shape = n, bias, activation, input.shape
        return acts[-1].view(*input.shape[:-1], weights[-1].shape[0])

    @staticmethod
    def backward(ctx, grad_o):
        n = ctx.n
        acts, weights = ctx.saved_tensors[:n + 1], ctx.saved_tensors[n + 1:]
        dy = grad_o.reshape(-1, grad_o.shape[-1]).contiguous()
        dws, dbs = [None] * n, [None] * n
        for i in range(n - 1, -1, -1):
            y = acts[i + 1]
            if ctx.activation == 1:
                dy = dy * (y > 0).to(dy.dtype)
            elif ctx.activation == 2:
                dy = (dy.float() * (y.float() * (1 - y.float()))).to(dy.dtype)
            dws[i] = G.linear_wgrad(dy, acts[i])
            if ctx.bias:
                dbs[i] = G.colsum(dy)
            if i > 0 or ctx.needs_input_grad[2]:
                dy = G.linear_dgrad(dy, weights[i].contiguous())
        dx = dy.view(ctx.in_shape) if ctx.needs_input_grad[2] else None
        return (None, None, dx, *dws, *(dbs if ctx.bias else []))


def mlp_function(bias, activation, input, *params):
    if torch.is_autocast_enabled():
        dt = torch.get_autocast_dtype("cuda")
        input = input.to(dt)
        params = tuple(p.to(dt) for p in params)
    with torch.amp.autocast("cuda", enabled=False):
        return MlpFunction.apply(bias, activation, input, *params)


class MLP(torch.nn.Module):
    def __init__(self, mlp_sizes, bias=True, activation="relu"):
        super().__init__()
        self.num_layers = len(mlp_sizes) - 1
        self.mlp_sizes = list(mlp_sizes)
        self.bias = 1 if bias else 0
        if activation not in _ACT:
            raise TypeError("activation must be relu or none.")
        self.activation = _ACT[activation]
        self.weights, self.biases = [], []
        for i in range(self.num_layers):
            w = nn.Parameter(torch.empty(mlp_sizes[i + 1], mlp_sizes[i]))
            self.weights.append(w)
            setattr(self, f"weight_{i}", w)
            if self.bias:
                b = nn.Parameter(torch.empty(mlp_sizes[i + 1]))
                self.biases.append(b)
                setattr(self, f"bias_{i}", b)
        self.reset_parameters()

    def reset_parameters(self):
        for w in self.weights:
            std = math.sqrt(2.0 / float(w.size(0) + w.size(1)))
            nn.init.normal_(w, 0.0, std)
        for b in self.biases:
            nn.init.normal_(b, 0.0, math.sqrt(1.0 / float(b.size(0))))

    def forward(self, input):
        return mlp_function(self.bias, self.activation, input, *self.weights, *self.biases)

    def extra_repr(self):
        return f"MLP sizes: {self.mlp_sizes}, Bias={self.bias}, activation={self.activation}"
