"""``LayerNormSmallShapeOptImpl.apply(inputs, normalized_shape, weight, bias, eps)`` — reference apex/contrib/openfold_triton/layer_norm.py:26
(Triton kernels tuned for OpenFold's many-rows / small-hidden shapes, incl. strided inputs)."""
from __future__ import annotations

import torch

from ...normalization.fused_layer_norm import fused_layer_norm_affine


class LayerNormSmallShapeOptImpl:
    @staticmethod
    def apply(inputs, normalized_shape, weight, bias, eps=1e-05):
        if not inputs.is_cuda:
            return torch.nn.functional.layer_norm(inputs, tuple(normalized_shape), weight, bias, eps)
        return fused_layer_norm_affine(inputs, weight, bias, tuple(normalized_shape), eps)
