"""``python -m apex_b200``: what is installed and how it is configured (library, kernels, device, run-time flags)."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

from . import __version__, _lib
from .utils import config


def _import_everything() -> None:
    """Signature declarations live next to each wrapper: import every sub-module so that the table is complete."""
    import importlib
    import pkgutil

    import apex_b200

    for m in pkgutil.walk_packages(apex_b200.__path__, "apex_b200."):
        if any(part in m.name for part in ("._C", "._build", "._kernels", ".csrc", "__main__")):
            continue
        try:
            importlib.import_module(m.name)
        except Exception:  # noqa: BLE001
            pass


def info() -> dict:
    _import_everything()
    pkg = Path(__file__).resolve().parent
    so = pkg / "_kernels.so"
    out = {"version": __version__, "package": str(pkg), "torch": torch.__version__, "cuda_runtime": torch.version.cuda,
           "kernels_library": str(so) if so.exists() else None, "host_runtime": None, "declared_entry_points": len(_lib._SIGS),
           "exported_entry_points": None, "cuda_available": torch.cuda.is_available(), "devices": [],
           "flags": config.flags()}
    _lib._try_load()
    if _lib._kernels is not None:
        out["host_runtime"] = getattr(_lib._C, "__file__", None)
        out["exported_entry_points"] = sum(hasattr(_lib._kernels, n) for n in _lib._SIGS)
    elif _lib._load_error is not None:
        out["load_error"] = repr(_lib._load_error)
    if torch.cuda.is_available():
        for i in range(torch.cuda.device_count()):
            p = torch.cuda.get_device_properties(i)
            out["devices"].append({"index": i, "name": p.name, "sm": f"{p.major}{p.minor}", "sms": p.multi_processor_count,
                                   "memory_gb": round(p.total_memory / 2 ** 30, 1)})
    return out


def main() -> int:
    d = info()
    if "--json" in sys.argv:
        print(json.dumps(d, indent=1))
        return 0
    print(f"apex_b200 {d['version']}  (torch {d['torch']}, CUDA runtime {d['cuda_runtime']})")
    print(f"  kernels : {d['kernels_library'] or 'NOT BUILT (python -m apex_b200._build)'}")
    if d.get("load_error"):
        print(f"  load error: {d['load_error']}")
    else:
        print(f"  entry points: {d['exported_entry_points']} exported / {d['declared_entry_points']} declared")
    if d["devices"]:
        for g in d["devices"]:
            note = "" if g["sm"] == "100" else "   <- kernels are built for sm_100a only"
            print(f"  cuda:{g['index']} {g['name']}  sm_{g['sm']}  {g['sms']} SMs  {g['memory_gb']} GB{note}")
    else:
        print("  no CUDA device: modules run on their PyTorch reference paths")
    print("  flags   : " + ", ".join(f"{k}={v}" for k, v in d["flags"].items()))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
