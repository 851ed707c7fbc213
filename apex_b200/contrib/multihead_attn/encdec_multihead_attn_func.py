"""Import path of the reference (apex/contrib/multihead_attn/encdec_multihead_attn_func.py); implementation in :mod:`.funcs`."""
from .funcs import EncdecAttnFunc, encdec_attn_func  # noqa: F401
