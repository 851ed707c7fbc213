"""CPU tests of the auxiliary subsystems: profiling helpers, config flags, deprecation warning, distributed test base (gloo)."""
import warnings

import pytest

import torch


def test_profiling_helpers_are_noops_without_cuda():
    from apex_b200.utils.profiling import ProfilerWindow, annotate, annotate_modules, nvtx_range
    with nvtx_range("x"):
        pass

    @annotate()
    def f(a):
        return a + 1

    assert f(1) == 2
    w = ProfilerWindow(2, 3)
    for i in range(8):
        w.step(i)
    m = torch.nn.Sequential(torch.nn.Linear(2, 2), torch.nn.ReLU())
    hs = annotate_modules(m)
    m(torch.randn(1, 2))
    assert len(hs) == 4
    for h in hs:
        h.remove()


def test_config_flags(monkeypatch):
    from apex_b200.utils import config
    monkeypatch.setenv("APEX_B200_DIST_NVLS", "0")
    assert config.dist_nvls_policy() == "off"
    monkeypatch.setenv("APEX_B200_DIST_NVLS", "1")
    assert config.dist_nvls_policy() == "on"
    monkeypatch.delenv("APEX_B200_DIST_NVLS")
    assert config.dist_nvls_policy() == "auto" and "APEX_B200_LN_FWD_V" in config.flags()


def test_deprecated_warning_and_apex_alias():
    import apex_b200
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        apex_b200.deprecated_warning("old thing")
    assert any(issubclass(x.category, apex_b200.DeprecatedFeatureWarning) for x in w)
    apex_b200.install_as_apex()
    import apex  # noqa: F401
    from apex.optimizers import FusedAdam  # noqa: F401
    from apex._autocast_utils import _cast_if_autocast_enabled
    assert _cast_if_autocast_enabled(1, 2) == (1, 2)


def test_reference_import_paths_resolve_to_the_same_objects():
    """One-class-per-file paths of the reference (``apex.optimizers.fused_novograd``, ``apex.contrib.optimizers.fp16_optimizer``,
    ``apex.contrib.sparsity.permutation_search_kernels.exhaustive_search`` ...) resolve through the alias finder to the grouped modules
    here, as the SAME objects (no second execution of any file)."""
    import apex_b200
    apex_b200.install_as_apex()
    from apex.contrib.optimizers.fp16_optimizer import FP16_Optimizer
    from apex.contrib.optimizers.fused_adam import FusedAdam as LegacyAdam
    from apex.contrib.sparsity.permutation_search_kernels import accelerated_search_for_good_permutation
    from apex.contrib.sparsity.permutation_search_kernels.call_permutation_search_kernels import accelerated_search_for_good_permutation as again
    from apex.contrib.sparsity.permutation_search_kernels.channel_swap import Channel_Swap  # noqa: F401
    from apex.contrib.sparsity.permutation_search_kernels.exhaustive_search import Exhaustive_Search  # noqa: F401
    from apex.multi_tensor_apply.multi_tensor_apply import MultiTensorApply
    from apex.optimizers.fused_adagrad import FusedAdagrad
    from apex.optimizers.fused_mixed_precision_lamb import FusedMixedPrecisionLamb
    from apex.optimizers.fused_novograd import FusedNovoGrad
    import apex.contrib.xentropy.softmax_xentropy as a
    import apex_b200.contrib.optimizers.legacy as legacy
    import apex_b200.contrib.xentropy.softmax_xentropy as b
    import apex_b200.multi_tensor_apply as mta
    import apex_b200.optimizers as O
    assert a is b and again is accelerated_search_for_good_permutation
    assert FusedNovoGrad is O.FusedNovoGrad and FusedAdagrad is O.FusedAdagrad and FusedMixedPrecisionLamb is O.FusedMixedPrecisionLamb
    assert FP16_Optimizer is legacy.FP16_Optimizer and LegacyAdam is legacy.FusedAdam and MultiTensorApply is mta.MultiTensorApply


def test_flatten_roundtrip():
    from apex_b200.utils.flatten import flatten, unflatten
    ts = [torch.randn(3, 4), torch.randn(5)]
    flat = flatten(ts)
    back = unflatten(flat, ts)
    assert all(torch.equal(a, b) for a, b in zip(ts, back))


class _Case:
    pass


def test_distributed_test_base_runs_ranks_on_gloo():
    import unittest

    from tests._dist_cases import GlooAllReduceCase
    suite = unittest.defaultTestLoader.loadTestsFromTestCase(GlooAllReduceCase)
    res = unittest.TextTestRunner(verbosity=0).run(suite)
    assert res.wasSuccessful(), res.failures + res.errors


def test_kernel_calls_run_under_a_device_guard_when_tensors_live_elsewhere(monkeypatch):
    """_lib.stream_ptr(device) arms a guard that _Fn.__call__ applies (multi-device processes; reference: OptionalCUDAGuard per op)."""
    import types

    from apex_b200 import _lib
    events = []

    class Guard:
        def __init__(self, d):
            self.d = d

        def __enter__(self):
            events.append(("enter", self.d))

        def __exit__(self, *a):
            events.append(("exit", self.d))

    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", Guard)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: types.SimpleNamespace(cuda_stream=77))
    f = _lib._Fn("probe", lambda *a: events.append(("call", a)) or 0)
    f(1, _lib.stream_ptr(torch.device("cuda", 1)))
    f(2, _lib.stream_ptr(torch.device("cuda", 0)))
    assert events == [("enter", 1), ("call", (1, 77)), ("exit", 1), ("call", (2, 77))]


def test_apex_alias_resolves_deep_imports_to_the_same_module_objects():
    import importlib

    import apex_b200

    apex_b200.install_as_apex()
    for name in ("contrib.xentropy.softmax_xentropy", "optimizers.fused_adam", "contrib.openfold_triton", "_lib", "transformer.functional.fused_softmax"):
        assert importlib.import_module("apex." + name) is importlib.import_module("apex_b200." + name), name
    from apex.contrib.openfold_triton.fused_adam_swa import FusedAdamSWA
    from apex_b200.contrib.openfold import FusedAdamSWA as same

    assert FusedAdamSWA is same
    # a package attribute that shadows a sub-module of the same name must survive a deep import of that sub-module under the alias
    for pkg in ("contrib.index_mul_2d", "contrib.focal_loss"):
        leaf = pkg.rsplit(".", 1)[1]
        fn = getattr(importlib.import_module("apex_b200." + pkg), leaf)
        assert callable(fn)
        assert importlib.import_module(f"apex.{pkg}.{leaf}") is importlib.import_module(f"apex_b200.{pkg}.{leaf}")
        assert getattr(importlib.import_module("apex_b200." + pkg), leaf) is fn and getattr(importlib.import_module("apex." + pkg), leaf) is fn
    with pytest.raises(ImportError):
        importlib.import_module("apex.no_such_module")


def test_info_cli_reports_the_build():
    from apex_b200.__main__ import info

    d = info()
    assert d["version"] and d["declared_entry_points"] >= 50 and "APEX_B200_DIST_NVLS" in d["flags"]
    if d["kernels_library"]:
        assert d["exported_entry_points"] == d["declared_entry_points"]


def test_dist_harness_fails_fast_with_the_failing_ranks_traceback():
    import time

    from apex_b200.testing.dist_harness import run_distributed
    from tests import _dist_cases as cases

    t0 = time.monotonic()
    with pytest.raises(RuntimeError, match="deliberate failure on rank 1"):
        run_distributed(cases.one_rank_raises_while_the_others_wait, 3, "cpu", backend="gloo", timeout=120.0)
    assert time.monotonic() - t0 < 60.0   # first failure + grace period, not the peers' collective time-out


def test_every_apex_b200_import_in_the_repo_resolves():
    """GPU-only scripts (benchmarks, examples, GPU tests, CUDA branches of the package) cannot be executed here, but their imports can be
    resolved: every ``from apex_b200... import X`` names something that exists, and an imported FUNCTION or CLASS is never used as if it were
    the sub-module of the same name (``from pkg import focal_loss as FL; FL.FocalLoss`` — the package re-exports a function called focal_loss)."""
    import ast
    import importlib
    import pathlib
    import types

    root = pathlib.Path(__file__).resolve().parent.parent
    files = [p for p in root.joinpath("apex_b200").rglob("*.py") if "csrc" not in p.parts]
    files += list(root.joinpath("tests").glob("*.py")) + list(root.joinpath("benchmarks").glob("*.py")) + list(root.joinpath("examples").rglob("*.py"))
    files += [root / "bench.py", root / "__graft_entry__.py"]
    problems = []
    for path in files:
        rel = path.relative_to(root)
        pkg = None
        if rel.parts[0] == "apex_b200":
            mod = ".".join(rel.with_suffix("").parts)
            pkg = mod[:-9] if mod.endswith(".__init__") else mod.rsplit(".", 1)[0]
        tree = ast.parse(path.read_text())
        resolved = {}
        for node in ast.walk(tree):
            if not isinstance(node, ast.ImportFrom):
                continue
            base = node.module or ""
            if node.level:
                if pkg is None:
                    continue
                parts = pkg.split(".")
                base = ".".join(parts[:len(parts) - (node.level - 1)] + ([node.module] if node.module else []))
            if not base.startswith("apex_b200") or base.startswith("apex_b200._C"):
                continue
            try:
                module = importlib.import_module(base)
            except Exception as e:  # noqa: BLE001
                problems.append(f"{rel}:{node.lineno}: import {base}: {e!r}")
                continue
            for a in node.names:
                if a.name == "*":
                    continue
                try:
                    obj = getattr(module, a.name)
                except AttributeError:
                    try:
                        obj = importlib.import_module(base + "." + a.name)
                    except Exception:  # noqa: BLE001
                        problems.append(f"{rel}:{node.lineno}: {base} has no {a.name}")
                        continue
                resolved.setdefault(a.asname or a.name, []).append(obj)
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in resolved:
                objs = resolved[node.value.id]
                if all(isinstance(o, (types.FunctionType, type)) and not hasattr(o, node.attr) for o in objs):
                    problems.append(f"{rel}:{node.lineno}: {node.value.id}.{node.attr} on a {type(objs[0]).__name__}")
    assert not problems, "\n".join(problems)
