"""OpenFold helper kernels. The reference package (apex/contrib/openfold_triton, 2.7k lines) is written in Triton with shipped
autotune tables for Ampere/Hopper; Triton is a compatibility layer this library does not use, so the same entry points run on the
sm_100a kernels: ``FusedAdamSWA`` on the multi-tensor engine, ``LayerNormSmallShapeOptImpl`` on the row-in-registers LayerNorm
(it adapts threads-per-row at run time, which is what the Triton autotune tables approximate), attention through the fused softmax
kernels. ``sync_triton_auto_tune_cache_across_gpus`` is a no-op (there is no JIT cache to broadcast)."""
from .fused_adam_swa import AdamMathType, FusedAdamSWA
from .layer_norm import LayerNormSmallShapeOptImpl
from .mha import AttnBiasJIT, AttnNoBiasJIT, AttnTri, CanSchTriMHA, FusedAttenionCoreFunc, disable, enable, is_enabled, schedule_triton_mha


def sync_triton_auto_tune_cache_across_gpus(strict: bool = True, verbose: bool = False) -> None:
    return None


__all__ = ("LayerNormSmallShapeOptImpl", "sync_triton_auto_tune_cache_across_gpus", "CanSchTriMHA", "AttnTri", "AttnBiasJIT",
           "AttnNoBiasJIT", "FusedAdamSWA", "AdamMathType", "enable", "disable", "is_enabled")
