"""In-tree build of the apex_b200 native code for sm_100a.

Two artefacts, both written next to this file so they travel with the source tree:
  * ``_kernels.so`` — every CUDA kernel + torch-free C-ABI launchers + the C++ symmetric-heap runtime (nvcc, sm_100a only).
  * ``_C.so``       — the torch/pybind host runtime (tensor tables), g++.

``python -m apex_b200._build`` (or ``__graft_entry__.build()``) compiles what is out of date; objects are cached under
``build/obj``. The reference builds 31 separate torch CUDAExtensions through setuptools (reference setup.py:24-1057); here the
kernels do not include torch headers, so a full rebuild is ~1-2 minutes on 8 cores.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
OBJ = ROOT / "build" / "obj"

NVCC = os.environ.get("NVCC", shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")
CUDA_HOME = Path(NVCC).resolve().parent.parent
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH_FLAGS + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                           "-Xcompiler", "-fvisibility=hidden", "-DNDEBUG"]


def _cutlass_include() -> list[str]:
    """CuTe/CUTLASS header tree vendored inside site-packages (used only as headers inside our own kernels)."""
    for base in sys.path:
        for rel in ("flashinfer/data/cutlass/include", "tilelang/3rdparty/cutlass/include"):
            p = Path(base) / rel
            if (p / "cute" / "tensor.hpp").exists():
                return ["-I", str(p)]
    return []


def _digest(paths: list[Path], extra: str) -> str:
    h = hashlib.sha1(extra.encode())
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout[-4000:], r.stderr[-8000:]))
    if r.stderr.strip() and os.environ.get("APEX_B200_BUILD_VERBOSE"):
        sys.stderr.write(r.stderr)


def _compile_cu(src: Path, headers: list[Path], verbose: bool) -> tuple[Path, bool]:
    flags = NVCC_FLAGS + ["-I", str(CSRC)]
    if "cutlass" in src.read_text()[:4000] or "cute/" in src.read_text()[:4000]:
        flags = flags + _cutlass_include()
    key = _digest([src] + headers, " ".join(flags))
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".sha1")
    if obj.exists() and stamp.exists() and stamp.read_text() == key:
        return obj, False
    if verbose:
        print(f"[apex_b200 build] nvcc {src.name}", flush=True)
    _run([NVCC] + flags + ["-c", str(src), "-o", str(obj)])
    stamp.write_text(key)
    return obj, True


def build_kernels(verbose: bool = True) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h"))
    srcs = sorted(CSRC.glob("*.cu")) + sorted(p for p in CSRC.glob("*.cpp") if p.name != "binding.cpp")
    out = PKG / "_kernels.so"
    jobs = max(1, min(len(srcs), int(os.environ.get("MAX_JOBS", os.cpu_count() or 4))))
    with ThreadPoolExecutor(jobs) as ex:
        res = list(ex.map(lambda s: _compile_cu(s, headers, verbose), srcs))
    objs = [o for o, _ in res]
    if any(ch for _, ch in res) or not out.exists():
        if verbose:
            print(f"[apex_b200 build] link {out.name}", flush=True)
        _run([NVCC] + ARCH_FLAGS + ["-shared", "-o", str(out)] + [str(o) for o in objs] + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"])
    return out


def build_binding(verbose: bool = True) -> Path:
    import torch
    from torch.utils.cpp_extension import include_paths

    OBJ.mkdir(parents=True, exist_ok=True)
    src = CSRC / "binding.cpp"
    out = PKG / "_C.so"
    tlib = Path(torch.__file__).resolve().parent / "lib"
    abi = int(torch.compiled_with_cxx11_abi())
    inc = []
    for p in include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"], "-isystem", str(CUDA_HOME / "include")]
    flags = ["-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_EXTENSION_NAME=_C",
             "-DTORCH_API_INCLUDE_EXTENSION_H", "-fvisibility=hidden", "-w"]
    key = _digest([src], " ".join(flags) + torch.__version__)
    stamp = OBJ / "binding.sha1"
    if out.exists() and stamp.exists() and stamp.read_text() == key:
        return out
    if verbose:
        print("[apex_b200 build] g++ binding.cpp", flush=True)
    cxx = os.environ.get("CXX", shutil.which("g++") or "g++")
    _run([cxx] + flags + inc + [str(src), "-o", str(out), f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
                                "-ltorch", "-ltorch_python", f"-Wl,-rpath,{tlib}"])
    stamp.write_text(key)
    return out


def build_all(verbose: bool = True) -> None:
    with ThreadPoolExecutor(2) as ex:
        a = ex.submit(build_kernels, verbose)
        b = ex.submit(build_binding, verbose)
        a.result()
        b.result()


if __name__ == "__main__":
    build_all(True)
    print("[apex_b200 build] ok")
