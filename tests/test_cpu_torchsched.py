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


# ---- generated multi-stream programs (contrib/torchsched/inductor) ---------------------------------------------------------------------
from apex_b200.contrib.torchsched import config as ts_config                                      # noqa: E402
from apex_b200.contrib.torchsched.inductor import _utils as ind_utils                             # noqa: E402
from apex_b200.contrib.torchsched.inductor import patch_graph_lowering                            # noqa: E402
from apex_b200.contrib.torchsched.inductor.event import CudaEventFactory                          # noqa: E402
from apex_b200.contrib.torchsched.inductor.scheduler import MultiCudaStreamScheduler              # noqa: E402


def test_generated_program_matches_eager_including_backward():
    torch.manual_seed(0)
    m = Branchy()
    x = torch.randn(32, 64, requires_grad=True)
    patch_graph_lowering(True)
    try:
        cm, graphs = _compile(m)
        out = cm(x)
    finally:
        patch_graph_lowering(False)
    ref = m(x)
    assert graphs[0].wrapper_codegen and "def call(" in graphs[0].program().source
    torch.testing.assert_close(out, ref)
    torch.testing.assert_close(torch.autograd.grad(out.sum(), x)[0], torch.autograd.grad(ref.sum(), x)[0])
    cm2, graphs2 = _compile(m)                       # not patched any more: the interpreter
    cm2(x)
    assert not graphs2[0].wrapper_codegen and graphs2[0]._program is None


class _SimStream:
    """Vector-clock model of a CUDA stream: ``clock[s]`` = how much of stream ``s``'s work is known to precede what this stream does next."""
    count = 0

    def __init__(self, device=None, name=None):
        _SimStream.count += 1
        self.name = name or f"side{_SimStream.count}"
        self.clock = {self.name: 0}
        self.log = []

    def tick(self):
        self.clock[self.name] += 1
        return self.clock[self.name]

    def knows(self, stream_name, t):
        return self.clock.get(stream_name, -1) >= t

    def wait_event(self, ev):
        assert ev.snapshot is not None, "wait on an event that was never recorded: a no-op on CUDA, i.e. a missed dependency"
        self.log.append("wait")
        for k, v in ev.snapshot.items():
            self.clock[k] = max(self.clock.get(k, -1), v)


class _SimEvent:
    created = 0

    def __init__(self):
        _SimEvent.created += 1
        self.snapshot = None

    def record(self, stream):
        stream.log.append("record")
        self.snapshot = dict(stream.clock)


def _simulate(monkeypatch, sched, args):
    """Run the generated program of ``sched`` under the stream model; every tensor an op reads must have been produced at a time the
    reading stream knows about. Returns (outputs, streams, number of events created)."""
    caller = _SimStream(name="caller")
    current = [caller]
    produced = {}                                    # id(tensor) -> (stream name, time)

    @contextlib.contextmanager
    def use(stream):
        current.append(stream)
        try:
            yield
        finally:
            current.pop()

    monkeypatch.setattr(torch.cuda, "Stream", _SimStream)
    monkeypatch.setattr(torch.cuda, "Event", _SimEvent)
    monkeypatch.setattr(torch.cuda, "stream", use)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: current[-1])
    monkeypatch.setattr(torch.cuda.nvtx, "range_push", lambda *_: None)
    monkeypatch.setattr(torch.cuda.nvtx, "range_pop", lambda *_: None)
    monkeypatch.setattr(ind_utils, "_pools", {})
    _SimEvent.created = 0
    sched.codegen()
    fn = sched.compile()

    def tracked(target):
        def run(*a, **k):
            s = current[-1]
            for t in S._tensors((a, k)):
                if id(t) in produced:
                    name, when, _ = produced[id(t)]
                    assert s.knows(name, when), f"{getattr(target, '__name__', target)} on {s.name} reads a value of {name} it never waited for"
            out = target(*a, **k)
            now = s.tick()
            for t in S._tensors(out):
                produced[id(t)] = (s.name, now, t)       # holding the tensor keeps its id from being recycled
            return out
        return run

    targets = fn.__globals__["_ts_t"]
    targets[:] = [tracked(t) for t in targets]
    now = caller.tick()
    for t in S._tensors(args):
        produced[id(t)] = ("caller", now, t)
    out = fn(*args)
    for t in S._tensors(out):                        # the caller's stream owns the results
        name, when, _ = produced.get(id(t), ("caller", 0, None))
        assert caller.knows(name, when), f"result produced on {name} is returned before the caller's stream waited for it"
    streams = {s for s in fn.__globals__.values() if isinstance(s, _SimStream)}
    for s in streams:                                # join: nothing a side stream did is left unordered
        assert caller.knows(s.name, s.clock[s.name])
    return out, streams, _SimEvent.created


def _annotate(gm, *args):
    """``example_value`` metadata (what dynamo attaches) for a hand-built graph."""
    class Rec(torch.fx.Interpreter):
        def run_node(self, n):
            out = super().run_node(n)
            n.meta["example_value"] = out
            return out
    Rec(gm).run(*args)
    return gm


def test_generated_program_stream_discipline(monkeypatch):
    torch.manual_seed(0)
    m = Branchy()
    x = torch.randn(16, 64)
    cm, graphs = _compile(m)
    ref = m(x)
    cm(x)
    sg = graphs[0]
    sched = MultiCudaStreamScheduler(sg.gm, multi_stream=True)
    out, streams, n_events = _simulate(monkeypatch, sched, sg.last_args)
    torch.testing.assert_close(out[0], ref)
    src = sched.wrapper.source
    assert len(streams) == 2 and n_events == 1 + 2                 # entrance event + one per branch
    assert src.count("with torch.cuda.stream(") == 2 and src.count(".wait_event(event0)") == 2
    assert src.count("default_stream.wait_event(") == 2 and "_ts_record_stream(y1, default_stream)" in src
    assert src.index("h = ") < src.index("with torch.cuda.stream(stream1)")      # critical path is issued first, on the caller's stream


def _random_graph(seed, n_ops=40):
    g = torch.Generator().manual_seed(seed)
    graph = torch.fx.Graph()
    root = nn.Module()
    root.w = nn.Parameter(torch.randn(64, 64, generator=g) / 8)
    vals = [graph.placeholder("x0"), graph.placeholder("x1")]
    w = graph.get_attr("w")
    unary, binary = [torch.tanh, torch.relu, torch.sigmoid], [torch.add, torch.mul, torch.sub]
    pick = lambda k: int(torch.randint(0, k, (1,), generator=g))                                  # noqa: E731
    for _ in range(n_ops):
        kind = pick(4)
        a = vals[max(0, len(vals) - 1 - pick(6))] if pick(2) else vals[pick(len(vals))]
        if kind == 0:
            vals.append(graph.call_function(torch.matmul, (a, w)))
        elif kind == 1:
            vals.append(graph.call_function(unary[pick(3)], (a,)))
        elif kind == 2:
            vals.append(graph.call_function(binary[pick(3)], (a, vals[pick(len(vals))])))
        else:
            vals.append(graph.call_method("mul", (a, 0.5)))
    leaves = [v for v in vals[2:] if not v.users]
    graph.output(tuple(leaves))
    return torch.fx.GraphModule(root, graph)


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("num_streams", [1, 3, 8])
def test_generated_programs_of_random_dags_are_race_free_and_exact(monkeypatch, seed, num_streams):
    gm = _random_graph(seed)
    args = (torch.randn(8, 64), torch.randn(8, 64))
    want = gm(*args)
    _annotate(gm, *args)
    sched = MultiCudaStreamScheduler(gm, num_streams=num_streams, multi_stream=True)
    out, streams, n_events = _simulate(monkeypatch, sched, args)
    for a, b in zip(out, want):
        torch.testing.assert_close(a, b)
    assert len(streams) <= num_streams
    # every wait in the text refers to an event recorded earlier in the text, and no stream waits twice for the same record
    lines = [ln.strip() for ln in sched.wrapper.source.splitlines()]
    live = {"event0"}
    for ln in lines:
        if ".record(" in ln:
            live.add(ln.split(".record(")[0])
        elif ".wait_event(" in ln:
            assert ln.split(".wait_event(")[1].split(")")[0] in live


def test_event_reuse_bounds_the_number_of_events(monkeypatch):
    gm = _random_graph(3, n_ops=60)
    args = (torch.randn(8, 64), torch.randn(8, 64))
    _annotate(gm, *args)
    with ts_config.patch(reuse_cuda_event=True):
        _, _, reused = _simulate(monkeypatch, MultiCudaStreamScheduler(gm, num_streams=4, multi_stream=True), args)
    with ts_config.patch(reuse_cuda_event=False):
        sched = MultiCudaStreamScheduler(gm, num_streams=4, multi_stream=True)
        _, _, fresh = _simulate(monkeypatch, sched, args)
    n_compute = sum(n.op in S._COMPUTE_OPS for n in gm.graph.nodes)
    assert reused < fresh <= n_compute + 1


def test_event_factory_lifecycle():
    f = CudaEventFactory(reuse_cuda_event=True)
    e0 = f.get_entrance_event()
    assert e0.materialized_event == "event0" and f.get_entrance_event() is e0
    a, b = f.get_sym_event(1), f.get_sym_event(2)
    assert a < b and a != b and len({a, b, f.get_sym_event(1)}) == 3 and "idx=1" in str(a)
    with pytest.raises(ValueError):
        a.wait(1)

    class Code:
        def __init__(self):
            self.lines = []

        def writeline(self, s):
            self.lines.append(s)

    code = Code()
    rec_a, wait_a, wait_a2, rec_b = a.record(1), a.wait(0), a.wait(2), b.record(2)
    rec_a.codegen(code)
    wait_a.codegen(code)
    assert a.materialized_event == "event1"
    wait_a2.codegen(code)
    assert a.materialized_event is None and f.available_materialized_events == ["event1"]
    rec_b.codegen(code)                                      # nobody waits for b: no line, no event
    assert code.lines == ["event1.record(stream1)", "default_stream.wait_event(event1)", "stream2.wait_event(event1)  # last wait of event 1"]
    c = f.get_sym_event(1)
    lines = [c.record(1), c.wait(0)]
    for ln in lines:
        ln.codegen(code)
    assert code.lines[-2:] == ["event1.record(stream1)", "default_stream.wait_event(event1)  # last wait of event 4"] and f.created == ["event1"]


def test_stream_pool_names_and_reuse(monkeypatch):
    monkeypatch.setattr(torch.cuda, "Stream", _SimStream)
    assert ind_utils.get_stream_name(0) == "default_stream" and ind_utils.get_stream_name(3) == "stream3"
    pool = ind_utils.CUDAStreamPool(pool_size=2)
    a = pool.acquire()
    b = pool.acquire()
    with pytest.raises(RuntimeError):
        pool.acquire()
    pool.release(a)
    assert pool.acquire() is a and pool.side_stream(2) is b
    with pytest.raises(IndexError):
        pool.side_stream(3)


def test_config_parsing_and_code_dump(tmp_path):
    assert ts_config._parse_ids("1,2,3-5,7-8") == {1, 2, 3, 4, 5, 7, 8} and ts_config._parse_ids("") == set()
    assert ts_config._parse_dump("+inductor,/tmp/x") == (["torchsched", "inductor"], "/tmp/x")
    assert ts_config._parse_dump("rel/dir")[0] == ["torchsched"] and ts_config._parse_dump("")[1] is None
    from apex_b200.contrib.torchsched.inductor import lower_graph
    gm = _random_graph(1, n_ops=6)
    args = (torch.randn(2, 64), torch.randn(2, 64))
    _annotate(gm, *args)
    with ts_config.patch(dump_code_dir=str(tmp_path), dump_code_backends=["torchsched", "inductor"]):
        fn = lower_graph(gm, graph_id=7, multi_stream=False)
    for a, b in zip(fn(*args), gm(*args)):
        torch.testing.assert_close(a, b)
    assert (tmp_path / "torchsched" / "graph_7_wrapper_code.py").read_text() == fn.source
    assert "def forward" in (tmp_path / "inductor" / "graph_7_wrapper_code.py").read_text()


# ---- backend: decompositions, AOT-autograd mode, pre-grad passes, graph-level LayerNorm ops ------------------------------------------------
from apex_b200.contrib.torchsched import backend as ts_backend                                    # noqa: E402
from apex_b200.contrib.torchsched.ops import layer_norm as ts_ln                                  # noqa: E402
from apex_b200.contrib.torchsched.passes import pre_grad_passes as ts_passes                      # noqa: E402


class ConvNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1, self.c2 = nn.Conv2d(4, 8, 3, padding=1), nn.Conv2d(8, 8, 3, padding=1, stride=2)
        self.ln, self.fc = nn.LayerNorm(8), nn.Linear(8, 4)

    def forward(self, x):
        h = F.relu(self.c1(x))
        return self.fc(self.ln(self.c2(h).permute(0, 2, 3, 1))).sum(-1)


@pytest.mark.parametrize("scheme", ["dwb", "wbd"])
def test_convolution_backward_decompositions_match_the_fused_call(monkeypatch, scheme):
    torch.manual_seed(0)
    x, w, go = torch.randn(2, 4, 9, 9), torch.randn(6, 2, 3, 3), torch.randn(2, 6, 5, 5)
    geometry = ([6], [2, 2], [1, 1], [1, 1], False, [0, 0], 2)
    decomp = getattr(ts_backend, f"convolution_backward_decomp_{scheme}")
    assert decomp(go, x, w, *geometry, [True, True, True]) is NotImplemented          # CPU tensors: nothing to overlap, stay fused
    monkeypatch.setattr(ts_backend, "_SPLIT_DEVICES", {"cuda", "cpu"})
    assert decomp(go, x, w, *geometry, [True, True, False]) is NotImplemented         # no bias gradient asked for: stay fused
    want = torch.ops.aten.convolution_backward(go, x, w, *geometry, [True, True, True])
    for got, ref in zip(decomp(go, x, w, *geometry, [True, True, True]), want):
        torch.testing.assert_close(got, ref)
    dx, dw, db = decomp(go, x, w, *geometry, [False, True, True])
    assert dx is None
    torch.testing.assert_close(dw, want[1])
    torch.testing.assert_close(db, want[2])


@pytest.mark.parametrize("scheme", ["dwb", "wbd"])
def test_aot_mode_schedules_forward_and_backward_graphs(monkeypatch, scheme):
    monkeypatch.setattr(ts_backend, "_SPLIT_DEVICES", {"cuda", "cpu"})
    torch._dynamo.reset()
    torch.manual_seed(0)
    m = ConvNet()
    x = torch.randn(2, 4, 8, 8, requires_grad=True)
    be = ts.get_backend("torchsched", scheme)
    assert isinstance(be, ts.DecompositionsWrapper) and be == ts.get_backend("torchsched", scheme) and be != ts.get_backend("torchsched", "dwb" if scheme == "wbd" else "wbd")
    with ts_config.patch(aot_autograd=True):
        out = torch.compile(m, backend=be)(x)
    ref = m(x)
    torch.testing.assert_close(out, ref)
    params = list(m.parameters())
    for g, r in zip(torch.autograd.grad(out.sum(), [x] + params), torch.autograd.grad(ref.sum(), [x] + params)):
        torch.testing.assert_close(g, r, atol=1e-4, rtol=1e-4)
    fw, bw = be.graphs
    assert bw.graph_id == fw.graph_id + 1
    fw_names, bw_names = [n.name for n in fw.plan.order], [n.name for n in bw.plan.order]
    assert "norm_fwd" in fw_names and "norm_bwd" in bw_names                         # LayerNorm is one fused node in each direction
    # each convolution's backward became separate weight / bias / data gradient nodes, in the scheme's order, on more than one stream
    convs = [n for n in bw.plan.order if n.target is torch.ops.aten.convolution_backward.default]
    sums = [n for n in bw.plan.order if n.target is torch.ops.aten.sum.dim_IntList and n.args[1] == [0, 2, 3]]
    assert len(convs) == 4 and len(sums) == 2
    masks = [tuple(n.args[-1]) for n in convs]
    want = [(True, False, False), (False, True, False)] if scheme == "dwb" else [(False, True, False), (True, False, False)]
    assert masks == want * 2
    assert len({bw.plan.stream_of[n] for n in convs + sums}) >= 3
    assert all(fw.plan.cost[n] == 0.0 for n in fw.plan.order if n.name.startswith(("detach", "view", "t")))   # views launch nothing


def test_aot_mode_with_generated_programs(monkeypatch):
    monkeypatch.setattr(ts_backend, "_SPLIT_DEVICES", {"cuda", "cpu"})
    torch._dynamo.reset()        # equal DecompositionsWrappers share dynamo's cache entry: start from a clean one
    torch.manual_seed(0)
    m = ConvNet()
    x = torch.randn(2, 4, 8, 8, requires_grad=True)
    be = ts.get_backend("torchsched")
    compile_fn = ts_backend.enable_multi_stream_scheduling(lambda: torch.compile(m, backend=be)(x))
    with ts_config.patch(aot_autograd=True):
        out = compile_fn()
    assert not ts_config.wrapper_codegen and all(sg.wrapper_codegen for sg in be.graphs)
    ref = m(x)
    torch.testing.assert_close(out, ref)
    torch.testing.assert_close(torch.autograd.grad(out.sum(), x)[0], torch.autograd.grad(ref.sum(), x)[0], atol=1e-4, rtol=1e-4)
    assert "convolution_backward" in be.graphs[1].program().source


def test_aot_backward_program_is_race_free(monkeypatch):
    monkeypatch.setattr(ts_backend, "_SPLIT_DEVICES", {"cuda", "cpu"})
    torch._dynamo.reset()
    torch.manual_seed(0)
    m = ConvNet()
    x = torch.randn(2, 4, 8, 8, requires_grad=True)
    be = ts.get_backend("torchsched")
    captured = {}
    orig = be._schedule

    def keep(gm, example_inputs=None, wrapper_codegen=None):
        sg = orig(gm, example_inputs)

        class Spy:
            def __call__(self, *a):
                captured[sg.graph_id] = a
                return sg(*a)
        return Spy()

    be._schedule = keep
    with ts_config.patch(aot_autograd=True):
        out = torch.compile(m, backend=be)(x)
    out.sum().backward()
    fw, bw = be.graphs
    want = bw(*captured[bw.graph_id])
    sched = MultiCudaStreamScheduler(bw.gm, multi_stream=True)
    got, streams, _ = _simulate(monkeypatch, sched, captured[bw.graph_id])
    assert len(streams) >= 3
    for a, b in zip(got, want):
        if a is not None:
            torch.testing.assert_close(a, b)


def test_get_backend_arguments():
    assert ts.get_backend("torch") == "inductor" and ts.get_backend("inductor") == "inductor"
    with pytest.raises(ValueError):
        ts.get_backend("tvm")
    with pytest.raises(ValueError):
        ts.get_backend("torchsched", "bdw")


def test_pre_grad_pass_normalises_arguments_and_counts():
    def f(x, w, b):
        return F.layer_norm(x, (16,), bias=b, weight=w) + F.layer_norm(x, (16,), w, b, 1e-3) + F.layer_norm(x, (16,))

    gm = torch.fx.symbolic_trace(f)
    calls = []

    def spy(x, normalized_shape, weight, bias, eps):
        calls.append((tuple(normalized_shape), weight is not None, bias is not None, eps))
        return F.layer_norm(x, normalized_shape, weight, bias, eps)

    assert ts_passes.run_pre_grad_pass("spy", gm.graph, F.layer_norm, spy) == 3
    x, w, b = torch.randn(4, 16), torch.randn(16), torch.randn(16)
    torch.testing.assert_close(gm(x, w, b), f(x, w, b))
    assert calls == [((16,), True, True, 1e-5), ((16,), True, True, 1e-3), ((16,), False, False, 1e-5)]
    assert ts_passes.run_pre_grad_pass("spy", gm.graph, F.layer_norm, spy) == 0
    with pytest.raises(ValueError):
        ts_passes.register_pattern("fused_layer_norm", F.layer_norm, spy)
    gm2 = torch.fx.symbolic_trace(f)
    before = ts_passes.counters.get("pre_grad_fused_layer_norm", 0)
    ts_passes.pre_grad_custom_pass(gm2.graph, traceable=True)
    assert ts_passes.counters["pre_grad_fused_layer_norm"] == before + 3
    assert sum(n.target is ts_passes.replace_layer_norm_traceable for n in gm2.graph.nodes) == 3
    with ts_config.patch(pre_grad_pass_options=["nope"]), pytest.raises(AssertionError):
        ts_passes.pre_grad_custom_pass(gm2.graph)
    torch.testing.assert_close(ts_passes.replace_layer_norm(x, (16,), w, b, 1e-5), F.layer_norm(x, (16,), w, b, 1e-5))


def test_graph_level_layer_norm_ops():
    torch.manual_seed(0)
    x = torch.randn(3, 5, 32, requires_grad=True)
    w, b = torch.randn(32, requires_grad=True), torch.randn(32, requires_grad=True)
    y, mean, invstd = ts_ln.layer_norm(x, [32], w, b, 1e-5)
    ref = F.layer_norm(x, (32,), w, b, 1e-5)
    torch.testing.assert_close(y, ref)
    torch.testing.assert_close(mean, x.detach().reshape(15, 32).mean(1))
    for got, want in zip(ts_ln.layer_norm_fake(x, [32], w, b), (y, mean, invstd)):
        assert got.shape == want.shape and got.dtype == want.dtype
    gy = torch.randn_like(y)
    dx, dw, db = ts_ln.layer_norm_backward(gy, mean, invstd, x.detach(), [32], w.detach(), b.detach())
    rx, rw, rb = torch.autograd.grad(ref, [x, w, b], gy)
    torch.testing.assert_close(dx, rx, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(dw, rw, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(db, rb, atol=1e-5, rtol=1e-5)
    for got, want in zip(ts_ln.layer_norm_backward_fake(gy, mean, invstd, x, [32], w, b), (dx, dw, db)):
        assert got.shape == want.shape
    ax, = torch.autograd.grad(y, x, gy)                                              # the op itself is differentiable
    torch.testing.assert_close(ax, rx, atol=1e-5, rtol=1e-5)
    with pytest.raises(ValueError):
        ts_ln.layer_norm(x, [16], w, b)


def test_transformer_layer_forward_and_backward_programs(monkeypatch):
    """A Hugging Face BERT layer through the AOT mode with generated programs: values and gradients equal eager, and every graph dynamo /
    AOT autograd produced (forward and backward fragments) is race-free under the stream model."""
    transformers = pytest.importorskip("transformers")
    from transformers.models.bert.modeling_bert import BertLayer
    cfg = transformers.BertConfig(hidden_size=64, num_attention_heads=4, intermediate_size=128, hidden_dropout_prob=0.0,
                                  attention_probs_dropout_prob=0.0)
    torch.manual_seed(0)
    torch._dynamo.reset()
    layer = BertLayer(cfg).eval()
    x = torch.randn(2, 10, 64, requires_grad=True)
    be = ts.get_backend("torchsched")
    captured = {}
    orig = be._schedule

    def keep(gm, example_inputs=None, wrapper_codegen=None):
        sg = orig(gm, example_inputs, wrapper_codegen)

        def run(*a):
            captured[sg.graph_id] = a
            return sg(*a)
        return run

    be._schedule = keep
    with ts_config.patch(aot_autograd=True, wrapper_codegen=True):
        out = torch.compile(layer, backend=be)(x)[0]
        (g,) = torch.autograd.grad(out.sum(), x)
    ref = layer(x)[0]
    torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(g, torch.autograd.grad(ref.sum(), x)[0], atol=1e-5, rtol=1e-5)
    assert len(be.graphs) >= 2 and all(sg.wrapper_codegen for sg in be.graphs)
    checked = 0
    for sg in be.graphs:
        if sg.plan.streams_used == 0 or sg.graph_id not in captured:
            continue
        args = captured[sg.graph_id]
        want = sg.gm(*args)
        got, streams, _ = _simulate(monkeypatch, MultiCudaStreamScheduler(sg.gm, multi_stream=True), args)
        assert len(streams) == sg.plan.streams_used
        for a, b in zip(got, want):
            if isinstance(a, torch.Tensor):
                torch.testing.assert_close(a, b)
        checked += 1
    assert checked >= 2


def test_codegen_fuzz_over_many_node_kinds():
    """tests/fuzz/fuzz_torchsched_codegen.py (modules, kwargs, getitem of multi-output ops, slices, dtype / inf constants, python scalars,
    tuple / dict / single outputs): the generated program returns exactly what the graph module returns; 300 seeds run offline."""
    import os
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz", "fuzz_torchsched_codegen.py")
    out = subprocess.run([sys.executable, script, "0", "40"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().splitlines()[-1] == "bad 0", out.stdout[-2000:] + out.stderr[-2000:]
