"""torchsched: the scheduling plan (critical path, stream assignment, event placement) is a pure function of the FX graph and is checked
here; stream execution is checked against a recording mock of the CUDA stream / event API (no GPU needed)."""
import contextlib

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from apex_b200.contrib import torchsched as ts
from apex_b200.contrib.torchsched import scheduler as S


class Branchy(nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.b, self.c, self.d = nn.Linear(64, 256), nn.Linear(256, 64), nn.Linear(64, 8), nn.Linear(64, 8)
        self.ln = nn.LayerNorm(64)

    def forward(self, x):
        h = self.b(F.gelu(self.a(x)))           # heavy chain: the critical path
        y1 = self.c(x).relu()                    # two light, independent branches
        y2 = torch.tanh(self.d(x))
        return self.ln(h + x).sum(-1, keepdim=True) + y1 + y2


def _compile(model, **kw):
    graphs = []

    def backend(gm, example_inputs):
        sg = ts._backend(gm, example_inputs, **kw)
        graphs.append(sg)

        def run(*args):          # this torch lifts parameters to graph inputs: keep what dynamo passes
            sg.last_args = args
            return sg(*args)

        return run

    return torch.compile(model, backend=backend), graphs


def test_plan_and_numerics_match_eager_including_backward():
    torch.manual_seed(0)
    m = Branchy()
    x = torch.randn(32, 64, requires_grad=True)
    cm, graphs = _compile(m)
    out, ref = cm(x), m(x)
    torch.testing.assert_close(out, ref)
    (g,) = torch.autograd.grad(out.sum(), x)
    (g_ref,) = torch.autograd.grad(ref.sum(), x)
    torch.testing.assert_close(g, g_ref)
    plan = graphs[0].plan
    names = [n.name for n in plan.critical_path]
    assert names[:3] == ["linear", "gelu", "h"] and all(plan.stream_of[n] == 0 for n in plan.critical_path)
    side = {plan.stream_of[n] for n in plan.order} - {0}
    assert len(side) == 2                                     # one side stream per independent branch
    by_name = {n.name: n for n in plan.order}
    assert plan.stream_of[by_name["linear_2"]] == plan.stream_of[by_name["y1"]] != plan.stream_of[by_name["y2"]]   # a chain stays on its stream
    assert {n.name for n in plan.records} == {"y1", "y2"}     # events only on cross-stream edges
    assert any(n.target is ts.fused_layer_norm_op for n in plan.order)   # pre-grad pass rewrote F.layer_norm
    assert "critical path" in plan.describe()


def test_in_place_graphs_and_zero_streams_stay_on_one_stream():
    class InPlace(nn.Module):
        def forward(self, x):
            y = x * 2
            y.add_(1)
            return y + torch.tanh(x)

    cm, graphs = _compile(InPlace())
    x = torch.randn(8)
    torch.testing.assert_close(cm(x), x * 2 + 1 + torch.tanh(x))
    assert graphs[0].plan.streams_used == 0 and "in-place" in graphs[0].plan.single_stream_reason
    cm, graphs = _compile(Branchy(), num_streams=0)
    cm(torch.randn(4, 64))
    assert graphs[0].plan.streams_used == 0


def test_registered_backend_name_runs_without_inductor():
    m = Branchy()
    x = torch.randn(4, 64)
    torch.testing.assert_close(torch.compile(m, backend="torchsched")(x), m(x))
    torch.testing.assert_close(ts.torchsched_compile(m)(x), m(x))
    assert "torchsched" in ts.list_backends()


class _FakeEvent:
    def __init__(self):
        self.recorded_on = None

    def record(self, stream):
        self.recorded_on = stream
        stream.log.append(("record", id(self)))


class _FakeStream:
    def __init__(self, name="side"):
        self.name, self.log = name, []

    def wait_event(self, ev):
        assert ev.recorded_on is not None, "waiting on an event that was never recorded is a no-op on CUDA: a missed dependency"
        assert ev.recorded_on is not self
        self.log.append(("wait", id(ev)))

    def wait_stream(self, other):
        self.log.append(("wait_stream", other.name))


def test_stream_execution_against_a_recording_mock(monkeypatch):
    torch.manual_seed(0)
    m = Branchy()
    x = torch.randn(16, 64)
    cm, graphs = _compile(m)
    ref = m(x)
    cm(x)                                                     # builds the plan on the plain CPU path
    sg = graphs[0]
    caller = _FakeStream("caller")
    current = [caller]

    @contextlib.contextmanager
    def use(stream):
        current.append(stream)
        try:
            yield
        finally:
            current.pop()

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "Stream", _FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", use)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: current[-1])
    monkeypatch.setattr(torch.cuda.nvtx, "range_push", lambda *_: None)
    monkeypatch.setattr(torch.cuda.nvtx, "range_pop", lambda *_: None)
    sg._streams = sg._events = None                           # re-create the resources from the mocked API
    out = sg(*sg.last_args)
    out = out[0] if isinstance(out, (tuple, list)) else out
    torch.testing.assert_close(out, ref)
    streams, events = sg._streams, sg._events
    assert len(streams) == 2 and len(events) == 2
    for s in streams:                                         # fork before any work, one record per branch, join at the end
        assert s.log[0] == ("wait_stream", "caller") and [op for op, _ in s.log].count("record") == 1
    assert [op for op, _ in caller.log].count("wait") == 2
    assert caller.log[-2:] == [("wait_stream", "side"), ("wait_stream", "side")]
