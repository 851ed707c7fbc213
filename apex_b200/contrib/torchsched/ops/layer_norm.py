"""``layer_norm`` op the pre-grad pass swaps in (reference apex/contrib/torchsched/ops/layer_norm.py:269-338 builds cuDNN graphs; here it is
the fused row-in-registers LayerNorm kernel of :mod:`apex_b200.normalization`)."""
from .. import fused_layer_norm_op


def layer_norm(x, normalized_shape, weight=None, bias=None, eps=1e-5):
    return fused_layer_norm_op(x, normalized_shape, weight, bias, eps)


__all__ = ["layer_norm"]
