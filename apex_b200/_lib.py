"""ctypes loader for ``_kernels.so`` (torch-free C ABI) and the ``_C`` pybind host runtime.

All kernel launchers return an int (0 = ok, >0 = cudaError_t, <0 = bad argument) and take the stream as the last argument.
Signatures are declared in ``_SIGS`` with a one-letter-per-argument code so ctypes converts/validates every call:
  p = pointer (int or None)   i = int32   l = int64   L = uint64   f = float   d = double
On a machine with a GPU a missing/broken native library is a hard error (no silent eager fallback); without a GPU the
pure-PyTorch reference implementations in ``apex_b200.ops.reference`` are used (CPU plumbing config of BASELINE.json).
"""
from __future__ import annotations

import ctypes
import threading
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_int64, "L": ctypes.c_uint64, "f": ctypes.c_float, "d": ctypes.c_double}

_SIGS: dict[str, str] = {}
_kernels = None
_C = None
_load_error: Exception | None = None

DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.float64: 3, torch.uint8: 4, torch.int32: 5,
      torch.int64: 6, torch.int16: 7}
if hasattr(torch, "float8_e4m3fn"):
    DT[torch.float8_e4m3fn] = 8
    DT[torch.float8_e5m2] = 9


def declare(name: str, spec: str) -> None:
    _SIGS[name] = spec.replace(" ", "")


def _try_load() -> None:
    global _kernels, _C, _load_error
    if _kernels is not None or _load_error is not None:
        return
    try:
        so = _PKG / "_kernels.so"
        if not so.exists():
            raise FileNotFoundError(f"{so} not built — run `python -m apex_b200._build`")
        _kernels = ctypes.CDLL(str(so), mode=ctypes.RTLD_GLOBAL)
        from . import _C as c  # noqa: PLC0415

        _C = c
    except Exception as e:  # noqa: BLE001
        _load_error = e
        _kernels = None
        _C = None


def available() -> bool:
    """True when the native library is loaded AND a CUDA device is usable."""
    _try_load()
    return _kernels is not None and torch.cuda.is_available()


def require() -> None:
    _try_load()
    if _kernels is None:
        raise RuntimeError(f"apex_b200 native library is not available: {_load_error!r}")


def host_runtime():
    """The pybind module (TensorTable etc.)."""
    require()
    return _C


_tls = threading.local()


class _Fn:
    __slots__ = ("name", "fn")

    def __init__(self, name: str, fn):
        self.name, self.fn = name, fn

    def __call__(self, *args):
        # Multi-device processes: every wrapper evaluates ``stream_ptr(tensor.device)`` among the arguments of this call; if that
        # device is not the current one the launch runs under a device guard (what at::cuda::OptionalCUDAGuard does in the reference).
        dev = getattr(_tls, "dev", None)
        if dev is not None:
            _tls.dev = None
            with torch.cuda.device(dev):
                rc = self.fn(*args)
        else:
            rc = self.fn(*args)
        if rc != 0:
            msg = f"cuda error {rc}" if rc > 0 else f"bad argument ({rc})"
            if rc > 0:
                try:
                    msg = torch.cuda.cudart().cudaGetErrorString(rc) if hasattr(torch.cuda.cudart(), "cudaGetErrorString") else msg
                except Exception:  # noqa: BLE001
                    pass
            raise RuntimeError(f"apex_b200 kernel {self.name} failed: {msg}")
        return rc


_cache: dict[str, _Fn] = {}


def fn(name: str) -> _Fn:
    f = _cache.get(name)
    if f is None:
        require()
        raw = getattr(_kernels, name)
        spec = _SIGS.get(name)
        if spec is None:
            raise KeyError(f"apex_b200: no signature declared for {name}")
        raw.argtypes = [_CT[c] for c in spec]
        raw.restype = ctypes.c_int
        f = _cache[name] = _Fn(name, raw)
    return f


def raw_fn(name: str, restype, argtypes):
    """For the few runtime entry points that do not follow the int-return convention."""
    require()
    raw = getattr(_kernels, name)
    raw.argtypes = argtypes
    raw.restype = restype
    return raw


def stream_ptr(device=None) -> int:
    """Raw handle of the current stream of ``device``; also arms the device guard of the kernel call being assembled (see _Fn)."""
    idx = getattr(device, "index", None)
    if idx is not None and idx != torch.cuda.current_device():
        _tls.dev = idx
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def dt(t: torch.Tensor | torch.dtype) -> int:
    d = t if isinstance(t, torch.dtype) else t.dtype
    try:
        return DT[d]
    except KeyError:
        raise TypeError(f"apex_b200: unsupported dtype {d}") from None


def gpu_required_error(what: str) -> RuntimeError:
    return RuntimeError(f"{what}: CUDA tensors given but the apex_b200 native library failed to load: {_load_error!r}")


# --- signature table (kept next to the loader so a mismatch is one grep away) ------------------------------------------
_T = "p i i i i"  # arena, n, depth, total_chunks, chunk
declare("ab_mt_scale", _T + " i i f p p p")
declare("ab_mt_axpby", _T + " i i i f f i p p")
declare("ab_mt_norm", _T + " i p p p p p p i i f f p")
declare("ab_mt_l2norm_scale", _T + " i i f p p p p p")
declare("ab_mt_adam", _T + " i i f f f f i i i f i p p p p p")
declare("ab_mt_adam_swa", _T + " i i f f f f i i i i f f f p p")
declare("ab_mt_adagrad", _T + " i f f i f p")
declare("ab_mt_sgd", _T + " i i i f f f f i i i f p p")
declare("ab_mt_novograd", _T + " i f f f f i i f i i p p")
declare("ab_update_scale_hysteresis", "p p p p d d i i p")
declare("ab_mt_lamb_stage1", _T + " i i i f f f i i f i f p p f p i p p p p p p p i p")
declare("ab_mt_lamb_stage2", _T + " i i p p f p f p i i p i i p")
declare("ab_mt_dist_adam", _T + " i i i p f f f f i i i f i p p p p")
declare("ab_mt_dist_adam_remainders", _T + " i p f f f f i i i f p")
declare("ab_mt_cast", _T + " i i f p")
