"""Import path of the reference (apex/contrib/multihead_attn/fast_self_multihead_attn_norm_add_func.py); implementation in :mod:`.funcs`."""
from .funcs import FastSelfAttnNormAddFunc, fast_self_attn_norm_add_func  # noqa: F401
