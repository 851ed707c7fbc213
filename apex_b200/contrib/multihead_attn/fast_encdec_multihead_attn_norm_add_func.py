"""Import path of the reference (apex/contrib/multihead_attn/fast_encdec_multihead_attn_norm_add_func.py); implementation in :mod:`.funcs`."""
from .funcs import FastEncdecAttnNormAddFunc, fast_encdec_attn_norm_add_func  # noqa: F401
