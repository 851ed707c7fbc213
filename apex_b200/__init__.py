"""apex_b200 — a Blackwell (B200, sm_100a) native mixed-precision / distributed training utility library with the
capabilities and public API of NVIDIA/apex.

Layout:
  csrc/            hand-written sm_100a CUDA kernels (torch-free C ABI) + C++ host runtime
  ops/             functional API over the kernels (amp_C-compatible multi-tensor ops, norms, softmax, GEMM, ...)
  optimizers/ normalization/ fused_dense/ mlp/ multi_tensor_apply/ transformer/ parallel/ contrib/   apex-compatible modules
  models/          parameter sets / small models used by the benchmarks
  utils/           device timers, roofline + clocks helpers
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from . import multi_tensor_apply, optimizers, normalization  # noqa: F401


class DeprecatedFeatureWarning(FutureWarning):
    pass


def deprecated_warning(msg: str) -> None:
    """Warn once per call site, on rank 0 only (reference apex/__init__.py:35-43)."""
    import warnings

    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_rank() == 0:
        warnings.warn(msg, DeprecatedFeatureWarning, stacklevel=2)


def install_as_apex() -> None:
    """Register ``apex`` / ``amp_C`` aliases in ``sys.modules`` so ``import apex`` user code runs on this library."""
    import sys

    from .ops import amp_C as _amp_C

    mod = sys.modules[__name__]
    sys.modules.setdefault("apex", mod)
    sys.modules.setdefault("amp_C", _amp_C)
    for name in ("optimizers", "normalization", "multi_tensor_apply", "fused_dense", "mlp", "parallel", "transformer", "contrib", "_autocast_utils",
                 "distributed_testing"):
        try:
            sub = __import__(f"{__name__}.{name}", fromlist=["*"])
            sys.modules.setdefault(f"apex.{name}", sub)
        except Exception:  # noqa: BLE001
            pass
    try:  # the reference's compiled-extension names (fused_layer_norm_cuda, scaled_*_softmax_cuda, xentropy_cuda, ...)
        from . import ext_compat

        ext_compat.install()
    except Exception:  # noqa: BLE001
        pass
