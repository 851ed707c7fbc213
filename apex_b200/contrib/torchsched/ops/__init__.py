"""Graph-level ops of the scheduler backend (reference apex/contrib/torchsched/ops/__init__.py imports the sub-module for its side effect
of registering the custom ops; here the ops are registered by apex_b200.normalization.custom_ops)."""
from . import layer_norm  # noqa: F401

__all__: list = []
