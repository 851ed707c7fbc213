"""Import path of the reference (apex/contrib/multihead_attn/fast_encdec_multihead_attn_func.py); implementation in :mod:`.funcs`."""
from .funcs import FastEncdecAttnFunc, fast_encdec_attn_func  # noqa: F401
