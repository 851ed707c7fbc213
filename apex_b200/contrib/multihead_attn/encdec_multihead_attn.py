"""Import path of the reference (apex/contrib/multihead_attn/encdec_multihead_attn.py)."""
from .funcs import jit_dropout_add  # noqa: F401
from .multihead_attn import EncdecMultiheadAttn  # noqa: F401
