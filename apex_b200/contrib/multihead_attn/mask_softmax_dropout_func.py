"""Import path of the reference (apex/contrib/multihead_attn/mask_softmax_dropout_func.py); implementation in :mod:`.funcs`."""
from .funcs import MaskSoftmaxDropout  # noqa: F401
from .multihead_attn import fast_mask_softmax_dropout_func  # noqa: F401
