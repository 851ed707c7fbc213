"""Program text of a scheduled graph: one Python function in which every FX node is one statement, wrapped in ``with torch.cuda.stream``
blocks and event record / wait lines. Reference: apex/contrib/torchsched/inductor/wrapper.py:39-288 (``MultiStreamWrapperCodegen`` and its
``Enter/ExitCudaStreamContextLine``), which extends Inductor's Python wrapper the same way for Inductor's fused kernels.

Why text and not the ``fx.Interpreter`` of :mod:`..scheduler`: the interpreter pays a dictionary lookup, an argument-tree map and a
context-manager round trip per node on every call; the generated function pays them once, at build time. The statements call exactly the
same targets (this library's kernels, ATen), so numerics are identical. The text is also what ``TORCH_SCHED_DUMP_CODE`` writes to disk."""
from __future__ import annotations

import dataclasses
import linecache
import math
import operator

import torch
import torch.fx as fx

from ._utils import DEFAULT_STREAM, DEFAULT_STREAM_IDX, ENTRANCE_EVENT, get_cuda_stream_pool, get_stream_name

__all__ = ["IndentedBuffer", "EnterCudaStreamContextLine", "ExitCudaStreamContextLine", "MultiStreamWrapperCodegen"]

_PREFIX = "_ts_"        # names of the generated function that are not FX node names


class IndentedBuffer:
    def __init__(self, indent: int = 0):
        self._lines: list[str] = []
        self._indent = indent

    def writeline(self, line: str) -> None:
        self._lines.append("    " * self._indent + line if line else "")

    def do_indent(self, n: int = 1) -> None:
        self._indent += n

    def do_unindent(self, n: int = 1) -> None:
        self._indent -= n
        assert self._indent >= 0

    def getvalue(self) -> str:
        return "\n".join(self._lines) + "\n"


@dataclasses.dataclass
class _Line:
    text: str

    def codegen(self, code: IndentedBuffer) -> None:
        code.writeline(self.text)


@dataclasses.dataclass
class EnterCudaStreamContextLine:
    """``with torch.cuda.stream(streamK):`` — the first time a side stream is entered it also waits for the entrance event, which
    orders it after everything the caller had queued before calling the graph."""
    stream_idx: int
    first_entry: bool

    def codegen(self, code: IndentedBuffer) -> None:
        name = get_stream_name(self.stream_idx)
        code.writeline(f"with torch.cuda.stream({name}):")
        code.do_indent()
        if self.first_entry:
            code.writeline(f"{name}.wait_event({ENTRANCE_EVENT})")


@dataclasses.dataclass
class ExitCudaStreamContextLine:
    def codegen(self, code: IndentedBuffer) -> None:
        code.do_unindent()


@dataclasses.dataclass
class EnterDeviceContextManagerWithStreamInfoLine:
    """Entry of a generated program (reference inductor/wrapper.py:39-66): name the caller's stream and record the entrance event on it —
    every side stream waits for that event the first time it is entered. The reference also emits the acquisition of
    ``config.num_streams`` pool streams here; in this design the streams are bound into the program's namespace once, at ``compile()``,
    so a call pays no pool traffic."""
    multi_stream: bool = True

    def codegen(self, code: IndentedBuffer) -> None:
        if self.multi_stream:
            code.writeline(f"{DEFAULT_STREAM} = torch.cuda.current_stream()")
            code.writeline(f"{ENTRANCE_EVENT}.record({DEFAULT_STREAM})")


@dataclasses.dataclass
class ExitDeviceContextManagerWithStreamInfoLine:
    """Exit of a generated program (reference inductor/wrapper.py:69-90 releases the pool streams): nothing to hand back here — the
    scheduler has already joined every side stream into the caller's stream through events, and the streams stay bound to the function."""

    def codegen(self, code: IndentedBuffer) -> None:
        return None


def record_stream_tree(value, stream) -> None:
    """``Tensor.record_stream`` on every CUDA tensor inside ``value``: the caching allocator must not hand a block that ``stream`` still
    reads to a later allocation of the stream that owns it."""
    if isinstance(value, torch.Tensor):
        if value.is_cuda:
            value.record_stream(stream)
    elif isinstance(value, (list, tuple)):
        for v in value:
            record_stream_tree(v, stream)
    elif isinstance(value, dict):
        for v in value.values():
            record_stream_tree(v, stream)


def _fetch_attr(root, target: str):
    for atom in target.split("."):
        root = getattr(root, atom)
    return root


class MultiStreamWrapperCodegen:
    """Collects the lines of one graph and turns them into a callable. The scheduler (:mod:`.scheduler`) decides WHAT is written; this
    class knows HOW each thing is spelled."""

    def __init__(self, gm: fx.GraphModule, graph_id: int = 0, multi_stream: bool = True) -> None:
        self.gm, self.graph_id, self.multi_stream = gm, graph_id, multi_stream
        self.lines: list = []
        self.targets: list = []       # call targets, referenced as _ts_t[i]
        self.constants: list = []     # argument values without a literal spelling, referenced as _ts_c[i]
        self.streams_entered: set[int] = set()
        self.max_stream_idx = 0
        for n in gm.graph.nodes:
            if n.name.startswith(_PREFIX) or n.name in ("torch", DEFAULT_STREAM):
                raise ValueError(f"graph value name {n.name!r} collides with the generated program's own names")

    # ---- spelling of values ---------------------------------------------------------------------------------------------------
    def _const(self, v) -> str:
        for i, c in enumerate(self.constants):
            if c is v:
                return f"{_PREFIX}c[{i}]"
        self.constants.append(v)
        return f"{_PREFIX}c[{len(self.constants) - 1}]"

    def fmt(self, a) -> str:
        if isinstance(a, fx.Node):
            return a.name
        if a is None or isinstance(a, (bool, int, str)) or a is Ellipsis:
            return repr(a)
        if isinstance(a, float):
            return repr(a) if math.isfinite(a) else f"float({str(a)!r})"
        if isinstance(a, tuple):
            inner = "".join(self.fmt(x) + ", " for x in a)
            return f"({inner})" if type(a) is tuple else f"{self._const(type(a))}({inner})"     # plain / named tuple
        if isinstance(a, list):
            return "[" + ", ".join(self.fmt(x) for x in a) + "]"
        if isinstance(a, dict):
            return "{" + ", ".join(f"{self.fmt(k)}: {self.fmt(v)}" for k, v in a.items()) + "}"
        if isinstance(a, slice):
            return f"slice({self.fmt(a.start)}, {self.fmt(a.stop)}, {self.fmt(a.step)})"
        return self._const(a)       # dtypes, devices, memory formats, named tuples, symbolic ints, ...

    def _call(self, head: str, args, kwargs) -> str:
        parts = [self.fmt(a) for a in args] + [f"{k}={self.fmt(v)}" for k, v in kwargs.items()]
        return f"{head}({', '.join(parts)})"

    # ---- lines ------------------------------------------------------------------------------------------------------------------
    def writeline(self, text: str) -> None:
        self.lines.append(_Line(text))

    def codegen_graph_nvtx_range_push(self) -> None:
        if self.multi_stream:
            self.writeline(f"torch.cuda.nvtx.range_push('torchsched graph {self.graph_id}')")

    def codegen_graph_nvtx_range_pop(self) -> None:
        if self.multi_stream:
            self.writeline("torch.cuda.nvtx.range_pop()")

    def codegen_device_guard_enter(self) -> None:
        """Entry of the program: name the caller's stream and record the entrance event on it."""
        self.lines.append(EnterDeviceContextManagerWithStreamInfoLine(self.multi_stream))

    def codegen_device_guard_exit(self) -> None:
        self.lines.append(ExitDeviceContextManagerWithStreamInfoLine())

    def codegen_cuda_stream_enter(self, stream_idx: int) -> None:
        assert stream_idx != DEFAULT_STREAM_IDX
        first = stream_idx not in self.streams_entered
        self.streams_entered.add(stream_idx)
        self.max_stream_idx = max(self.max_stream_idx, stream_idx)
        self.lines.append(EnterCudaStreamContextLine(stream_idx, first))

    def codegen_cuda_stream_exit(self) -> None:
        self.lines.append(ExitCudaStreamContextLine())

    def codegen_events_wait_stream(self, events, stream_idx: int) -> None:
        for ev in sorted(events):
            self.lines.append(ev.wait(stream_idx))

    def codegen_event_record(self, event, stream_idx: int) -> None:
        self.lines.append(event.record(stream_idx))

    def codegen_buffers_record_stream(self, names, stream_idx: int) -> None:
        for name in names:
            self.writeline(f"{_PREFIX}record_stream({name}, {get_stream_name(stream_idx)})")

    def codegen_free(self, names) -> None:
        if names:
            self.writeline("del " + ", ".join(names))

    def codegen_node(self, node: fx.Node) -> None:
        if node.op == "placeholder":
            return
        if node.op == "get_attr":
            self.writeline(f"{node.name} = {_PREFIX}getattr({_PREFIX}gm, {node.target!r})")
        elif node.op == "call_function":
            if node.target is operator.getitem and not node.kwargs:
                self.writeline(f"{node.name} = {self.fmt(node.args[0])}[{self.fmt(node.args[1])}]")
            else:
                self.targets.append(node.target)
                self.writeline(f"{node.name} = " + self._call(f"{_PREFIX}t[{len(self.targets) - 1}]", node.args, node.kwargs)
                               + f"  # {getattr(node.target, '__name__', node.target)}")
        elif node.op == "call_method":
            self.writeline(f"{node.name} = " + self._call(f"{self.fmt(node.args[0])}.{node.target}", node.args[1:], node.kwargs))
        elif node.op == "call_module":
            self.targets.append(self.gm.get_submodule(node.target))
            self.writeline(f"{node.name} = " + self._call(f"{_PREFIX}t[{len(self.targets) - 1}]", node.args, node.kwargs) + f"  # {node.target}")
        elif node.op == "output":
            self.writeline(f"{_PREFIX}out = {self.fmt(node.args[0])}")
        else:
            raise NotImplementedError(node.op)

    # ---- the program ------------------------------------------------------------------------------------------------------------
    def generate(self, event_factory=None) -> str:
        body = IndentedBuffer(indent=1)
        placeholders = [n.name for n in self.gm.graph.nodes if n.op == "placeholder"]
        if placeholders:
            body.writeline(", ".join(placeholders) + ("," if len(placeholders) == 1 else "") + f" = {_PREFIX}args")
        for line in self.lines:
            line.codegen(body)
        body.writeline(f"return {_PREFIX}out")
        head = IndentedBuffer()
        head.writeline(f"# torchsched graph {self.graph_id}: " + (f"{len(self.streams_entered)} side streams, "
                       f"{len(event_factory.created) if event_factory else 0} events" if self.multi_stream else "single stream"))
        head.writeline(f"def call(*{_PREFIX}args):")
        self.source = head.getvalue() + body.getvalue()
        self.event_names = list(event_factory.created) if (event_factory and self.multi_stream) else []
        return self.source

    def compile(self):
        """-> ``call(*args)``. Streams come from the process-wide pool, events are created once and live as long as the function."""
        ns = {"torch": torch, f"{_PREFIX}t": self.targets, f"{_PREFIX}c": self.constants, f"{_PREFIX}gm": self.gm,
              f"{_PREFIX}getattr": _fetch_attr, f"{_PREFIX}record_stream": record_stream_tree}
        if self.multi_stream:
            pool = get_cuda_stream_pool(pool_size=max(32, self.max_stream_idx))
            for s in sorted(self.streams_entered):
                ns[get_stream_name(s)] = pool.side_stream(s)
            for name in [ENTRANCE_EVENT] + self.event_names:
                ns[name] = torch.cuda.Event()
        filename = f"<torchsched graph {self.graph_id}>"
        linecache.cache[filename] = (len(self.source), None, self.source.splitlines(True), filename)
        exec(compile(self.source, filename, "exec"), ns)
        fn = ns["call"]
        fn.source = self.source
        return fn
