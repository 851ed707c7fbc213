"""``fused_weight_gradient_mlp_cuda`` equivalents: main_grad += dy^T @ x in ONE tcgen05 GEMM whose epilogue reads and adds the
persistent main-grad tile (beta = 1). Reference: csrc/megatron/fused_weight_gradient_dense_cuda.cu:17-83 (cublasGemmEx, beta=1)
and fused_weight_gradient_dense_16bit_prec_cuda.cu:17-75. Leading dimensions of x / dy are collapsed like the reference (:59-76)."""
from __future__ import annotations

import torch

from ...ops import gemm as G


def _collapse(t):
    return t.reshape(-1, t.shape[-1]).contiguous()


def wgrad_gemm_accum_fp32(input: torch.Tensor, d_output: torch.Tensor, main_grad: torch.Tensor) -> None:
    """main_grad (fp32 [out, in]) += d_output^T @ input, inputs fp32/fp16/bf16."""
    assert main_grad.dtype == torch.float32
    x, dy = _collapse(input), _collapse(d_output)
    if x.dtype == torch.float32:
        main_grad.addmm_(dy.t(), x)
        return
    G.linear_wgrad(dy, x, accum_into=main_grad)


def wgrad_gemm_accum_fp16(input: torch.Tensor, d_output: torch.Tensor, main_grad: torch.Tensor) -> None:
    """main_grad (fp16/bf16 [out, in]) += d_output^T @ input."""
    x, dy = _collapse(input), _collapse(d_output)
    assert main_grad.dtype == x.dtype
    G.linear_wgrad(dy, x, accum_into=main_grad)
