from .group_norm import GroupNorm, group_norm_nhwc, torch_group_norm

__all__ = ["GroupNorm", "group_norm_nhwc", "torch_group_norm"]
