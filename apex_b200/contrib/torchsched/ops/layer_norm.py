"""``layer_norm`` op the pre-grad pass swaps in (reference apex/contrib/torchsched/ops/layer_norm.py:269-338 builds cuDNN graphs; here it is
the fused row-in-registers LayerNorm kernel of :mod:`apex_b200.normalization`)."""
from .. import fused_layer_norm_op as layer_norm  # noqa: F401

__all__ = ["layer_norm"]
