"""Import path of the reference (apex/contrib/multihead_attn/self_multihead_attn_func.py); implementation in :mod:`.funcs`."""
from .funcs import SelfAttnFunc, self_attn_func  # noqa: F401
