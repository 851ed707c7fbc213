from .distributed_fused_adam import DistributedFusedAdam
from .distributed_fused_lamb import DistributedFusedLAMB
from .legacy import FP16_Optimizer, FusedAdam, FusedLAMB, FusedSGD

__all__ = ["DistributedFusedAdam", "DistributedFusedLAMB", "FusedAdam", "FusedLAMB", "FusedSGD", "FP16_Optimizer"]
