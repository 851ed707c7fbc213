from .distributed_fused_adam import DistributedFusedAdam

__all__ = ["DistributedFusedAdam"]
try:
    from .distributed_fused_lamb import DistributedFusedLAMB  # noqa: F401

    __all__.append("DistributedFusedLAMB")
except ImportError:
    pass
