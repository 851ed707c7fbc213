"""Import path of the reference (apex/contrib/multihead_attn/fast_self_multihead_attn_func.py); implementation in :mod:`.funcs`."""
from .funcs import FastSelfAttnFunc, fast_self_attn_func  # noqa: F401
