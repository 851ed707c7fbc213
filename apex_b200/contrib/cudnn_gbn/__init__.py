from .batch_norm import GroupBatchNorm2d

__all__ = ["GroupBatchNorm2d"]
