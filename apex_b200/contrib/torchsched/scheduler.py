"""Multi-stream scheduling of a captured FX graph — no code generation, no Inductor.

What the reference does (apex/contrib/torchsched/inductor/scheduler.py, graph.py, wrapper.py, event.py): after Inductor has fused the
graph it finds the longest (critical) path through the fused-node DAG, keeps that path on the default stream, round-robins the other
nodes over ``TORCH_SCHED_NUM_STREAMS`` side streams and emits wrapper code with ref-counted CUDA events on every cross-stream edge.

Here the same policy is applied to the graph dynamo hands to a backend, and the graph is *interpreted*: every node is an ordinary
eager call (so this library's hand-written kernels are what runs, and autograd records the ops — the backward of each op is then
executed by the autograd engine on the stream its forward ran on, with the engine's own cross-stream synchronisation). The plan —
critical path, stream per node, which nodes record an event, which edges wait — is a pure function of the graph and is what the CPU
tests check; on a machine without CUDA the interpreter simply runs the nodes in order.

A launch-bound graph can additionally be captured once into a CUDA graph (``ScheduledGraph(..., cuda_graph=True)``; inference only):
the side-stream forks and joins become graph dependencies and the per-node Python overhead disappears."""
from __future__ import annotations

import operator
from dataclasses import dataclass, field

import torch
import torch.fx as fx

from . import config

_COMPUTE_OPS = ("call_function", "call_method", "call_module")
# measured B200 numbers (MEASURED_PEAKS.json): only their ratio matters for ranking paths
_HBM_BYTES_PER_S, _TENSOR_FLOPS_PER_S, _LAUNCH_S = 6.5e12, 1.4e15, 3e-6
_MATMUL_NAMES = ("matmul", "linear", "bmm", "mm", "addmm", "baddbmm", "conv1d", "conv2d", "conv3d", "convolution", "einsum",
                 "scaled_dot_product_attention")


def _example(node: fx.Node):
    return node.meta.get("example_value", node.meta.get("val"))


def _tensors(v):
    if isinstance(v, torch.Tensor):
        yield v
    elif isinstance(v, (list, tuple)):
        for x in v:
            yield from _tensors(x)
    elif isinstance(v, dict):
        for x in v.values():
            yield from _tensors(x)


def _nbytes(v) -> int:
    return sum(t.numel() * t.element_size() for t in _tensors(v))


def _target_name(node: fx.Node) -> str:
    t = node.target
    return t if isinstance(t, str) else getattr(t, "__name__", str(t))


_VIEW_NAMES = frozenset(("view", "view_as", "t", "transpose", "permute", "expand", "expand_as", "unsqueeze", "squeeze", "detach", "narrow",
                         "select", "unbind", "chunk", "split", "unflatten", "alias"))


def _is_view(node: fx.Node) -> bool:
    """Metadata-only ops launch nothing: they cost nothing and stay on their producer's stream."""
    if isinstance(node.target, torch._ops.OpOverload):
        return bool(node.target.is_view)
    return node.op in ("call_method", "call_function") and _target_name(node) in _VIEW_NAMES


def estimate_cost(node: fx.Node, modules: dict | None = None) -> float:
    """Seconds, from the fake-tensor metadata: max(bytes moved / HBM rate, matmul FLOPs / tensor rate) + one launch. Nodes without
    tensor outputs (shape arithmetic, getitem) are free."""
    if node.op not in _COMPUTE_OPS:
        return 0.0
    out = _example(node)
    out_bytes = _nbytes(out)
    if out_bytes == 0 or node.target is operator.getitem or _is_view(node):
        return 0.0
    in_bytes = sum(_nbytes(_example(a)) for a in node.all_input_nodes)
    name = _target_name(node)
    if node.op == "call_module" and modules is not None:
        name = type(modules.get(node.target, None)).__name__.lower()
    flops = 0.0
    if any(k in name for k in _MATMUL_NAMES):
        # 2 * (output elements) * (reduction length); the reduction length is the last dim of the first tensor input
        first = next(_tensors(_example(node.all_input_nodes[0])), None) if node.all_input_nodes else None
        out_elems = sum(t.numel() for t in _tensors(out))
        if first is not None and first.dim() > 0:
            flops = 2.0 * out_elems * first.shape[-1]
    return max((in_bytes + out_bytes) / _HBM_BYTES_PER_S, flops / _TENSOR_FLOPS_PER_S) + _LAUNCH_S


def _mutates(node: fx.Node) -> bool:
    """In-place writes create write-after-read hazards that data-flow edges do not describe: such graphs stay on one stream."""
    if node.op == "call_method" and isinstance(node.target, str) and node.target.endswith("_") and not node.target.endswith("__"):
        return True
    if node.op == "call_function":
        if isinstance(node.target, torch._ops.OpOverload):      # ATen-level graphs (AOT autograd): the schema says it
            return node.target._schema.is_mutable
        if node.target in (operator.setitem, operator.iadd, operator.isub, operator.imul, operator.itruediv):
            return True
        if "out" in node.kwargs or _target_name(node).endswith("_"):
            return True
    return False


@dataclass
class Plan:
    order: list = field(default_factory=list)            # compute nodes in execution (= graph) order
    cost: dict = field(default_factory=dict)             # node -> estimated seconds
    critical_path: list = field(default_factory=list)    # longest path, source to sink
    stream_of: dict = field(default_factory=dict)        # node -> 0 (caller's stream) or 1..num_streams
    waits: dict = field(default_factory=dict)            # node -> producers on OTHER streams it must wait for
    records: set = field(default_factory=set)            # nodes that record an event after running
    single_stream_reason: str | None = None

    @property
    def streams_used(self) -> int:
        return len({s for s in self.stream_of.values() if s})

    def describe(self) -> str:
        lines = [f"torchsched plan: {len(self.order)} nodes, critical path {len(self.critical_path)} nodes "
                 f"({sum(self.cost[n] for n in self.critical_path) * 1e6:.1f} us of {sum(self.cost.values()) * 1e6:.1f} us), "
                 f"{self.streams_used} side streams" + (f" [{self.single_stream_reason}]" if self.single_stream_reason else "")]
        for n in self.order:
            w = ",".join(m.name for m in self.waits.get(n, ()))
            lines.append(f"  s{self.stream_of[n]} {n.name:<28} {self.cost[n] * 1e6:8.2f} us" + (f"  waits[{w}]" if w else "") + ("  records" if n in self.records else ""))
        return "\n".join(lines)


def plan_graph(gm: fx.GraphModule, num_streams: int | None = None) -> Plan:
    num_streams = config.num_streams if num_streams is None else num_streams
    modules = dict(gm.named_modules())
    plan = Plan()
    nodes = [n for n in gm.graph.nodes if n.op in _COMPUTE_OPS]
    plan.order = nodes
    plan.cost = {n: estimate_cost(n, modules) for n in nodes}
    if not nodes:
        return plan
    # longest path (by cost) ending at each node; graph order is a topological order
    best, prev = {}, {}
    for n in nodes:
        preds = [m for m in n.all_input_nodes if m in best]
        p = max(preds, key=lambda m: best[m], default=None)
        best[n] = plan.cost[n] + (best[p] if p is not None else 0.0)
        prev[n] = p
    tail = max(nodes, key=lambda n: best[n])
    while tail is not None:
        plan.critical_path.append(tail)
        tail = prev[tail]
    plan.critical_path.reverse()
    on_critical = set(plan.critical_path)

    if num_streams <= 0:
        plan.single_stream_reason = "TORCH_SCHED_NUM_STREAMS=0"
    elif any(_mutates(n) for n in nodes):
        plan.single_stream_reason = "graph contains in-place ops"
    if plan.single_stream_reason:
        plan.stream_of = {n: 0 for n in nodes}
        return plan

    rr = 0
    for n in nodes:
        if n in on_critical or plan.cost[n] == 0.0:
            # free nodes (views of python scalars, getitem) follow their producer so they never add a cross-stream edge
            src = next((m for m in n.all_input_nodes if m in plan.stream_of), None)
            plan.stream_of[n] = 0 if n in on_critical or src is None else plan.stream_of[src]
            continue
        side = [m for m in n.all_input_nodes if plan.stream_of.get(m, 0) != 0]
        if side:   # continue the chain of the most expensive side-stream producer: no new event on that edge
            plan.stream_of[n] = plan.stream_of[max(side, key=lambda m: plan.cost[m])]
        else:
            plan.stream_of[n] = 1 + rr % num_streams
            rr += 1
    for n in nodes:
        w = [m for m in n.all_input_nodes if m in plan.stream_of and plan.stream_of[m] != plan.stream_of[n] and _nbytes(_example(m)) > 0]
        if w:
            plan.waits[n] = w
            plan.records.update(w)
    return plan


class _Runner(fx.Interpreter):
    """fx.Interpreter whose ``run_node`` executes each node on its planned stream with event waits / records on cross-stream edges."""

    def __init__(self, gm: fx.GraphModule, plan: Plan, streams, events):
        super().__init__(gm, garbage_collect_values=True)
        self.plan, self.streams, self.events = plan, streams, events

    def run_node(self, n: fx.Node):
        if not self.streams or n.op not in _COMPUTE_OPS:
            return super().run_node(n)
        s = self.plan.stream_of.get(n, 0)
        stream = self.caller_stream if s == 0 else self.streams[s - 1]
        for m in self.plan.waits.get(n, ()):
            stream.wait_event(self.events[m])
            for t in _tensors(self.env.get(m)):   # the caching allocator must not recycle the block while this stream still reads it
                if t.is_cuda:
                    t.record_stream(stream)
        with torch.cuda.stream(stream):
            out = super().run_node(n)
        if n in self.plan.records:
            self.events[n].record(stream)
        return out

    def run(self, *args):
        if self.streams:
            self.caller_stream = torch.cuda.current_stream()
            used = sorted({s for s in self.plan.stream_of.values() if s})
            for s in used:                          # fork: side streams start after everything already queued by the caller
                self.streams[s - 1].wait_stream(self.caller_stream)
            out = super().run(*args)
            for s in used:                          # join: the caller's stream owns the results again
                self.caller_stream.wait_stream(self.streams[s - 1])
            for t in _tensors(out):
                if t.is_cuda:
                    t.record_stream(self.caller_stream)
            return out
        return super().run(*args)


class ScheduledGraph:
    """Callable replacement for ``gm.forward`` (what a dynamo backend returns)."""

    _next_id = 0

    def __init__(self, gm: fx.GraphModule, num_streams: int | None = None, cuda_graph: bool = False):
        self.gm = gm
        self.graph_id = ScheduledGraph._next_id
        ScheduledGraph._next_id += 1
        skip = self.graph_id in config.skip_graph_ids or self.graph_id in config.skip_post_grad_graph_ids
        self.num_streams = 0 if skip else num_streams
        self.plan = plan_graph(gm, self.num_streams)
        self.cuda_graph = cuda_graph
        self._streams = self._events = None
        self._captured = None
        # inductor.patch_graph_lowering(True) / TORCH_SCHED_CODEGEN=1 at construction time: run the generated program (inductor/), not
        # the interpreter; built on first use
        self.wrapper_codegen = bool(config.wrapper_codegen)
        self._program = None
        if config.debug:
            print(f"[torchsched] graph {self.graph_id}\n{self.plan.describe()}")

    def _resources(self):
        if self._streams is None:
            multi = torch.cuda.is_available() and self.plan.streams_used > 0
            n = max(self.plan.stream_of.values(), default=0)
            self._streams = [torch.cuda.Stream() for _ in range(n)] if multi else []
            self._events = {m: torch.cuda.Event() for m in self.plan.records} if multi else {}
        return self._streams, self._events

    def program(self):
        """The generated multi-stream program of this graph (``.source`` holds its text)."""
        if self._program is None:
            from .inductor.graph import lower_graph

            self._program = lower_graph(self.gm, self.num_streams, self.graph_id)
        return self._program

    def _run(self, *args):
        if self.wrapper_codegen:
            return self.program()(*args)
        streams, events = self._resources()
        if streams and not config.reuse_cuda_event:
            events = {m: torch.cuda.Event() for m in self.plan.records}
        if streams:
            torch.cuda.nvtx.range_push(f"graph {self.graph_id}")
        try:
            return _Runner(self.gm, self.plan, streams, events).run(*args)
        finally:
            if streams:
                torch.cuda.nvtx.range_pop()

    def __call__(self, *args):
        if not (self.cuda_graph and torch.cuda.is_available()) or torch.is_grad_enabled() and any(
                isinstance(a, torch.Tensor) and a.requires_grad for a in args):
            return self._run(*args)
        if self._captured is None:   # static-input capture: later calls copy into the captured buffers and replay
            static = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(2):
                    self._run(*static)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                out = self._run(*static)
            self._captured = (g, static, out)
        g, static, out = self._captured
        for dst, src in zip(static, args):
            if isinstance(dst, torch.Tensor):
                dst.copy_(src)
        g.replay()
        return out
