"""Symbolic CUDA events of the scheduling phase and their materialisation into the few real events a generated program creates.
Reference: apex/contrib/torchsched/inductor/event.py:30-206 (``CudaEventSym`` / ``CudaEventFactory``).

Scheduling hands every node that has a consumer on another stream a *symbolic* event; each ``wait`` taken on it bumps its reference
count. When the program text is emitted, a symbolic event that nobody waits for produces no code at all, and the others borrow a real
event (``event<k>``) from the factory for the span between their record line and their last wait line — after that the real event goes
back to the factory and a later record can take it again. Re-recording a CUDA event does not disturb waits that were already issued
(``cudaStreamWaitEvent`` captures the record that precedes it in program order), and the program is issued by one host thread, so
this is safe; it keeps the number of live events at the width of the graph instead of its node count
(``TORCH_SCHED_REUSE_CUDA_EVENT=0`` turns it off)."""
from __future__ import annotations

import dataclasses
import functools
import itertools

from .. import config as torchsched_config
from ._utils import DEFAULT_STREAM_IDX, ENTRANCE_EVENT, EVENT_NAME_TEMPLATE, get_stream_name

__all__ = ["CudaEventSym", "CudaEventFactory"]


@functools.total_ordering
@dataclasses.dataclass(eq=False)
class CudaEventSym:
    """Event number ``idx`` (program order) recorded on stream ``originate_stream_idx``."""
    factory: "CudaEventFactory"
    idx: int
    originate_stream_idx: int
    ref_count: int = 0
    materialized_event: str | None = None

    def _key(self):
        return (self.idx, self.originate_stream_idx)

    def __lt__(self, other):
        if not isinstance(other, CudaEventSym) or other.factory is not self.factory:
            return NotImplemented
        return self._key() < other._key()

    def __eq__(self, other):
        if not isinstance(other, CudaEventSym):
            return NotImplemented
        return other.factory is self.factory and self._key() == other._key()

    def __hash__(self):
        return hash((id(self.factory),) + self._key())

    def __str__(self):
        extra = (f", ref_count={self.ref_count}" if self.ref_count else "") + (
            f", materialized to `{self.materialized_event}`" if self.materialized_event else "")
        return f"CudaEventSym(idx={self.idx}, originate_stream_idx={self.originate_stream_idx}{extra})"

    def record(self, stream_idx: int) -> "_CudaEventRecordLine":
        """Program line recording this event on ``stream_idx``; emitted only if somebody waits for the event."""
        return _CudaEventRecordLine(self, get_stream_name(stream_idx))

    def wait(self, stream_idx: int) -> "_CudaEventWaitLine":
        """Program line making ``stream_idx`` wait for this event; takes a reference."""
        if stream_idx == self.originate_stream_idx:
            raise ValueError(f"stream {stream_idx} waiting for its own event {self}")
        self.ref_count += 1
        return _CudaEventWaitLine(self, get_stream_name(stream_idx))


@dataclasses.dataclass
class _CudaEventRecordLine:
    event: CudaEventSym
    stream: str

    def codegen(self, code) -> None:
        ev = self.event
        assert ev.materialized_event is None, f"{ev} recorded twice"
        if ev.ref_count > 0 or not ev.factory.reuse_cuda_event:
            ev.materialized_event = ev.factory.get_materialized_event(code)
            code.writeline(f"{ev.materialized_event}.record({self.stream})")


@dataclasses.dataclass
class _CudaEventWaitLine:
    event: CudaEventSym
    stream: str

    def codegen(self, code) -> None:
        ev = self.event
        assert ev.ref_count > 0 and ev.materialized_event is not None, f"wait emitted before the record of {ev}"
        line = f"{self.stream}.wait_event({ev.materialized_event})"
        ev.ref_count -= 1
        if ev.ref_count == 0 and ev is not ev.factory._entrance_event:
            ev.factory.deposit_materialized_event(ev.materialized_event)
            line += f"  # last wait of event {ev.idx}"
            ev.materialized_event = None
        code.writeline(line)


class CudaEventFactory:
    """Hands out symbolic events with increasing indices and lends real event names to them at emission time."""

    def __init__(self, reuse_cuda_event: bool | None = None) -> None:
        self.reuse_cuda_event = torchsched_config.reuse_cuda_event if reuse_cuda_event is None else reuse_cuda_event
        self._sym_idx = itertools.count(1)
        self._real_idx = itertools.count(1)
        self.available_materialized_events: list[str] = []
        self.created: list[str] = []          # every real event the program needs, in creation order
        self._entrance_event: CudaEventSym | None = None

    def get_entrance_event(self) -> CudaEventSym:
        """``event0``: recorded on the caller's stream before anything else; side streams wait for it before their first node."""
        if self._entrance_event is None:
            self._entrance_event = CudaEventSym(self, 0, DEFAULT_STREAM_IDX, materialized_event=ENTRANCE_EVENT)
        return self._entrance_event

    def get_sym_event(self, originate_stream_idx: int) -> CudaEventSym:
        return CudaEventSym(self, next(self._sym_idx), originate_stream_idx)

    def get_materialized_event(self, code=None) -> str:
        if self.reuse_cuda_event and self.available_materialized_events:
            return self.available_materialized_events.pop(0)
        name = EVENT_NAME_TEMPLATE.format(event_idx=next(self._real_idx))
        self.created.append(name)
        return name

    def deposit_materialized_event(self, event: str) -> None:
        assert event not in self.available_materialized_events, f"{event} returned twice"
        if self.reuse_cuda_event:
            self.available_materialized_events.append(event)
