"""CPU tests of the auxiliary subsystems: profiling helpers, config flags, deprecation warning, distributed test base (gloo)."""
import warnings

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
