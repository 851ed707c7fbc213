from .fused_adam import FusedAdam
from .fused_lamb import FusedLAMB, FusedMixedPrecisionLamb
from .fused_sgd import FusedAdagrad, FusedNovoGrad, FusedSGD

__all__ = ["FusedAdam", "FusedLAMB", "FusedMixedPrecisionLamb", "FusedSGD", "FusedNovoGrad", "FusedAdagrad"]
