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


def check_cudnn_version_and_warn(global_option: str, required_cudnn_version: int) -> bool:
    """True when cuDNN >= ``required_cudnn_version`` is usable, else warn and return False (reference apex/__init__.py:21-30)."""
    import warnings

    import torch

    ok = torch.backends.cudnn.is_available()
    version = torch.backends.cudnn.version() if ok else None
    if not (ok and version >= required_cudnn_version):
        warnings.warn(f"`{global_option}` depends on cuDNN {required_cudnn_version} or later, but "
                      f"{'cuDNN is not available' if not ok else version}")
        return False
    return True


class DeprecatedFeatureWarning(FutureWarning):
    pass


def deprecated_warning(msg: str) -> None:
    """Warn once per call site, on rank 0 only (reference apex/__init__.py:35-43)."""
    import warnings

    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_rank() == 0:
        warnings.warn(msg, DeprecatedFeatureWarning, stacklevel=2)


# Reference sub-module paths whose content lives in a differently named module here (one file per class in the reference, grouped files
# here). Resolved by the alias finder, so ``from apex.optimizers.fused_novograd import FusedNovoGrad`` works without a file per name.
_PATH_ALIASES = {
    "optimizers.fused_novograd": "optimizers.fused_sgd",
    "optimizers.fused_adagrad": "optimizers.fused_sgd",
    "optimizers.fused_mixed_precision_lamb": "optimizers.fused_lamb",
    "multi_tensor_apply.multi_tensor_apply": "multi_tensor_apply",
    "contrib.optimizers.fp16_optimizer": "contrib.optimizers.legacy",
    "contrib.optimizers.fused_adam": "contrib.optimizers.legacy",
    "contrib.optimizers.fused_lamb": "contrib.optimizers.legacy",
    "contrib.optimizers.fused_sgd": "contrib.optimizers.legacy",
    **{f"contrib.multihead_attn.{name}": "contrib.multihead_attn.funcs" for name in (
        "self_multihead_attn", "encdec_multihead_attn", "self_multihead_attn_func", "encdec_multihead_attn_func", "fast_self_multihead_attn_func",
        "fast_encdec_multihead_attn_func", "fast_self_multihead_attn_norm_add_func", "fast_encdec_multihead_attn_norm_add_func",
        "mask_softmax_dropout_func")},
    "contrib.sparsity.permutation_search_kernels": "contrib.sparsity.permutation_search",
    "contrib.sparsity.permutation_search_kernels.call_permutation_search_kernels": "contrib.sparsity.permutation_search",
    "contrib.sparsity.permutation_search_kernels.exhaustive_search": "contrib.sparsity.permutation_search",
    "contrib.sparsity.permutation_search_kernels.channel_swap": "contrib.sparsity.permutation_search",
    "contrib.sparsity.permutation_search_kernels.permutation_utilities": "contrib.sparsity.permutation_search",
}


def _pin_shadowed_submodules() -> None:
    """A package attribute may shadow a sub-module of the same name (``contrib.index_mul_2d.index_mul_2d`` is the function, as in the
    reference). Python binds ``package.child = module`` whenever a sub-module is LOADED, so loading such a sub-module a second time under
    its ``apex.`` name would replace the function by the module. Registering the alias in ``sys.modules`` up front means the import
    system finds it loaded and rebinds nothing."""
    import sys

    for name, m in list(sys.modules.items()):
        if m is None or not name.startswith(__name__ + "."):
            continue
        parent, _, child = name.rpartition(".")
        pm = sys.modules.get(parent)
        if pm is not None and getattr(pm, child, m) is not m:
            sys.modules.setdefault("apex" + name[len(__name__):], m)


class _ApexAliasFinder:
    """Meta-path finder that resolves ``apex.<anything>`` to the SAME module object as ``apex_b200.<anything>`` — without it a deep import
    such as ``apex.contrib.xentropy.softmax_xentropy`` would execute the file a second time under the other name (two copies of every class)."""

    def find_spec(self, fullname, path=None, target=None):
        import importlib
        import importlib.machinery
        import sys

        if not (fullname == "apex" or fullname.startswith("apex.")) or sys.modules.get("apex") is not sys.modules[__name__]:
            return None
        target_name = __name__ + "." + _PATH_ALIASES[fullname[5:]] if fullname[5:] in _PATH_ALIASES else __name__ + fullname[4:]
        try:
            importlib.import_module(target_name)
        except ImportError:
            return None
        _pin_shadowed_submodules()          # importing the target may have loaded sub-modules that a package attribute shadows
        spec = importlib.machinery.ModuleSpec(fullname, self)
        spec.loader_state = target_name
        return spec

    def create_module(self, spec):
        import sys
        import types

        target = sys.modules[spec.loader_state]
        if any(k.startswith(spec.name[5:] + ".") for k in _PATH_ALIASES) and not hasattr(target, "__path__"):
            # a reference PACKAGE that is a single module here (permutation_search_kernels): a package-shaped view of it, so that its
            # sub-module names keep resolving through this finder
            pkg = types.ModuleType(spec.name, target.__doc__)
            pkg.__dict__.update({k: v for k, v in target.__dict__.items() if not k.startswith("__")})
            pkg.__path__ = []
            return pkg
        return target

    def exec_module(self, module):  # already executed under its apex_b200 name
        return None


def install_as_apex() -> None:
    """Make ``import apex`` / ``import amp_C`` / ``import fused_layer_norm_cuda`` ... user code run on this library: ``apex`` and every
    ``apex.*`` sub-module become aliases of the corresponding ``apex_b200`` module objects (no effect if the real apex is already imported)."""
    import sys

    from .ops import amp_C as _amp_C

    mod = sys.modules[__name__]
    sys.modules.setdefault("apex", mod)
    sys.modules.setdefault("amp_C", _amp_C)
    if sys.modules["apex"] is mod and not any(isinstance(f, _ApexAliasFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _ApexAliasFinder())
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
