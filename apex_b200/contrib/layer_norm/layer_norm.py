"""``FastLayerNorm`` (reference apex/contrib/layer_norm/layer_norm.py:46-59 over 2,108 lines of per-hidden-size template
instantiations, ln_api.cpp:47-149). Here it is the same row-in-registers kernel as FusedLayerNorm — it already adapts
threads-per-row to the hidden size at run time (8..512 threads, 16-byte vectors), so no table of 26 compiled sizes is needed."""
from __future__ import annotations

import torch
from torch.nn import init

from ...normalization.fused_layer_norm import _NormFunction


class FastLayerNormFN:
    @staticmethod
    def apply(x, gamma, beta, epsilon, memory_efficient=False):
        hidden = gamma.numel()
        y = _NormFunction.apply(x.reshape(-1, hidden), gamma, beta, (hidden,), epsilon, memory_efficient, False, False)
        return y.view(x.shape)


def _fast_layer_norm(x, weight, bias, epsilon, memory_efficient):
    return FastLayerNormFN.apply(x, weight, bias, epsilon, memory_efficient)


class FastLayerNorm(torch.nn.Module):
    def __init__(self, hidden_size, eps=1e-5, memory_efficient=False):
        super().__init__()
        self.epsilon = eps
        self.memory_efficient = memory_efficient
        self.weight = torch.nn.Parameter(torch.empty(hidden_size))
        self.bias = torch.nn.Parameter(torch.empty(hidden_size))
        self.reset_parameters()

    def reset_parameters(self):
        init.ones_(self.weight)
        init.zeros_(self.bias)

    def forward(self, x):
        if not x.is_cuda:
            return torch.nn.functional.layer_norm(x, self.weight.shape, self.weight, self.bias, self.epsilon)
        return _fast_layer_norm(x, self.weight, self.bias, self.epsilon, self.memory_efficient)
