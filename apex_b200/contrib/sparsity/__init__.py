from .asp import ASP
from .permutation_lib import Permutation
from .sparse_masklib import create_mask

__all__ = ["ASP", "Permutation", "create_mask"]
