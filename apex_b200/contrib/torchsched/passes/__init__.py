"""Graph passes of the scheduler backend (reference apex/contrib/torchsched/passes/__init__.py)."""
from .pre_grad_passes import pre_grad_custom_pass

__all__ = ["pre_grad_custom_pass"]
