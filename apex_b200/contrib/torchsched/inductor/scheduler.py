"""Multi-stream scheduling of a graph into program text. Reference: apex/contrib/torchsched/inductor/scheduler.py:39-531
(``MultiCudaStreamScheduler``: critical path on the default stream, the rest round-robin over side streams, symbolic events on the
cross-stream edges, stream-context switching while the wrapper code is generated).

The reference schedules Inductor's fused nodes; here the nodes are those of the FX graph dynamo (or AOT autograd, see
:mod:`..backend`) produced, and the statements call this library's kernels / ATen directly. The stream assignment itself is
:func:`..scheduler.plan_graph` — shared with the interpreter — so both execution modes follow the same plan.

What this class adds on top of the plan:

* one symbolic event per producer that has a consumer on another stream (``register_downstream_event``);
* per consumer, the minimal set of events to wait for (``get_cross_stream_dependencies``): events of one stream complete in order, so a
  stream that already waited for event *k* of stream *s* needs no wait for an earlier event of *s*, and of several needed events of
  *s* only the last one is kept;
* the join at the end (``get_final_events_to_sync``): the caller's stream waits once per side stream, for that stream's last event,
  unless an earlier wait already covers it;
* ``record_stream`` on every buffer that crosses streams, and ``del`` of every buffer after its last use."""
from __future__ import annotations

import torch
import torch.fx as fx

from .. import config as torchsched_config
from ..scheduler import _COMPUTE_OPS, _example, _nbytes, plan_graph
from ._utils import DEFAULT_STREAM_IDX, get_stream_name
from .event import CudaEventFactory, CudaEventSym
from .wrapper import MultiStreamWrapperCodegen

__all__ = ["MultiCudaStreamScheduler"]


class MultiCudaStreamScheduler:
    def __init__(self, gm: fx.GraphModule, num_streams: int | None = None, graph_id: int = 0, multi_stream: bool | None = None) -> None:
        self.gm, self.graph_id = gm, graph_id
        self.plan = plan_graph(gm, num_streams)
        if multi_stream is None:
            multi_stream = torch.cuda.is_available()
        self.multi_stream = bool(multi_stream) and self.plan.streams_used > 0
        self.event_factory = CudaEventFactory()
        self.wrapper = MultiStreamWrapperCodegen(gm, graph_id, self.multi_stream)
        self.nodes = list(gm.graph.nodes)
        self.stream_idx_of: dict = {}
        self.node_event: dict = {}                 # producer -> its symbolic event
        self.last_event_of_stream: dict = {}       # stream -> latest event recorded on it
        self._waited: dict = {}                    # (waiting stream, recording stream) -> highest event index already waited for
        self._current_stream_idx: int | None = None
        self._recorded_on_current: set = set()
        self.schedule_multi_cuda_streams()

    # ---- state ------------------------------------------------------------------------------------------------------------------
    @property
    def current_stream_idx(self) -> int | None:
        return self._current_stream_idx

    @property
    def current_stream_name(self) -> str | None:
        return None if self._current_stream_idx is None else get_stream_name(self._current_stream_idx)

    @property
    def buffers_recorded_on_current_stream(self) -> set:
        """Names of foreign buffers already ``record_stream``-ed on the stream being written (cleared on every switch)."""
        return self._recorded_on_current

    @buffers_recorded_on_current_stream.setter
    def buffers_recorded_on_current_stream(self, names) -> None:
        self._recorded_on_current = set(names)

    def debug_str_short(self, node: fx.Node) -> str:
        return f"{node.name}@{get_stream_name(self.stream_idx_of.get(node, 0))} cost={self.plan.cost.get(node, 0.0) * 1e6:.2f}us"

    @staticmethod
    def get_last_event(events) -> CudaEventSym:
        return max(events)

    # ---- scheduling ---------------------------------------------------------------------------------------------------------------
    def schedule_multi_cuda_streams(self) -> None:
        """Stream index of every node. Compute nodes follow the plan; ``get_attr`` and ``output`` stay on the caller's stream."""
        for n in self.nodes:
            self.stream_idx_of[n] = self.plan.stream_of.get(n, DEFAULT_STREAM_IDX) if self.multi_stream else DEFAULT_STREAM_IDX

    def _crosses(self, producer: fx.Node, consumer: fx.Node) -> bool:
        return (self.stream_idx_of[producer] != self.stream_idx_of[consumer] and producer.op in _COMPUTE_OPS
                and _nbytes(_example(producer)) > 0)

    def register_downstream_event(self, node: fx.Node) -> CudaEventSym | None:
        """After ``node`` ran: give it an event if a consumer lives on another stream (the output node lives on the caller's)."""
        if not any(self._crosses(node, u) for u in node.users):
            return None
        s = self.stream_idx_of[node]
        ev = self.event_factory.get_sym_event(s)
        self.node_event[node] = self.last_event_of_stream[s] = ev
        self.wrapper.codegen_event_record(ev, s)
        return ev

    def get_cross_stream_dependencies(self, node: fx.Node):
        """(events the node's stream must wait for, producers whose buffers cross into it) — after removing what earlier waits cover."""
        t = self.stream_idx_of[node]
        newest: dict = {}
        foreign = []
        for m in node.all_input_nodes:
            if not self._crosses(m, node):
                continue
            foreign.append(m)
            ev = self.node_event[m]
            s = ev.originate_stream_idx
            if s not in newest or newest[s] < ev:
                newest[s] = ev
        events = set()
        for s, ev in newest.items():
            if self._waited.get((t, s), -1) >= ev.idx:
                continue
            self._waited[(t, s)] = ev.idx
            events.add(ev)
        return events, foreign

    def get_final_events_to_sync(self):
        """Join: the last event of every side stream that the caller's stream has not already waited for."""
        out = set()
        for s in sorted(self.wrapper.streams_entered):
            ev = self.last_event_of_stream.get(s)
            if ev is not None and self._waited.get((DEFAULT_STREAM_IDX, s), -1) < ev.idx:
                self._waited[(DEFAULT_STREAM_IDX, s)] = ev.idx
                out.add(ev)
        return out

    # ---- stream contexts ------------------------------------------------------------------------------------------------------------
    def generate_stream_ctx_enter(self, node: fx.Node) -> None:
        s = self.stream_idx_of[node]
        if s != DEFAULT_STREAM_IDX:
            self.wrapper.codegen_cuda_stream_enter(s)
        self._current_stream_idx = s
        self._recorded_on_current = set()

    def generate_stream_ctx_exit(self) -> None:
        if self._current_stream_idx not in (None, DEFAULT_STREAM_IDX):
            self.wrapper.codegen_cuda_stream_exit()
        self._current_stream_idx = None

    def generate_stream_ctx_switching(self, node: fx.Node) -> None:
        if self.stream_idx_of[node] != self._current_stream_idx:
            self.generate_stream_ctx_exit()
            self.generate_stream_ctx_enter(node)

    def propagate_cross_stream_dependencies(self, node: fx.Node) -> None:
        events, foreign = self.get_cross_stream_dependencies(node)
        t = self.stream_idx_of[node]
        self.wrapper.codegen_events_wait_stream(events, t)
        fresh = [m.name for m in foreign if m.name not in self._recorded_on_current]
        self.wrapper.codegen_buffers_record_stream(fresh, t)
        self._recorded_on_current.update(fresh)

    # ---- program ------------------------------------------------------------------------------------------------------------------
    def _find_tail_nodes(self) -> None:
        """A side stream whose last node has no cross-stream consumer (its value is dead, or consumed on the same stream) still has to be
        joined: give its last node an event."""
        last_node: dict = {}
        for n in self.nodes:
            s = self.stream_idx_of[n]
            if s != DEFAULT_STREAM_IDX and n.op in _COMPUTE_OPS:
                last_node[s] = n
        self._tail_nodes = {n for n in last_node.values()}

    def codegen(self) -> str:
        w = self.wrapper
        w.codegen_graph_nvtx_range_push()
        w.codegen_device_guard_enter()
        last_use: dict = {}
        for n in self.nodes:
            for m in n.all_input_nodes:
                last_use[m] = n
        frees: dict = {}
        for m, n in last_use.items():
            if n.op != "output" and m.op != "placeholder":
                frees.setdefault(n, []).append(m.name)
        self._find_tail_nodes()
        for n in self.nodes:
            if n.op == "placeholder":
                continue
            if n.op == "output":
                self.generate_stream_ctx_exit()
                self._current_stream_idx = DEFAULT_STREAM_IDX
                self.propagate_cross_stream_dependencies(n)
                w.codegen_events_wait_stream(self.get_final_events_to_sync(), DEFAULT_STREAM_IDX)
                w.codegen_node(n)
                break
            self.generate_stream_ctx_switching(n)
            if self.multi_stream:
                self.propagate_cross_stream_dependencies(n)
            w.codegen_node(n)
            if self.multi_stream and n.op in _COMPUTE_OPS:
                ev = self.register_downstream_event(n)
                if ev is None and n in self._tail_nodes:
                    s = self.stream_idx_of[n]
                    ev = self.event_factory.get_sym_event(s)
                    self.node_event[n] = self.last_event_of_stream[s] = ev
                    w.codegen_event_record(ev, s)
            w.codegen_free(frees.get(n, ()))
        w.codegen_device_guard_exit()
        w.codegen_graph_nvtx_range_pop()
        src = w.generate(self.event_factory)
        if torchsched_config.debug:
            print(src)
        return src

    def compile(self):
        if not hasattr(self.wrapper, "source"):
            self.codegen()
        return self.wrapper.compile()
