"""Functional API over the sm_100a kernels."""
from . import amp_C, reference  # noqa: F401
