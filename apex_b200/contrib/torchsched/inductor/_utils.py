"""Names and the stream pool shared by the generated multi-stream programs. Reference: apex/contrib/torchsched/inductor/_utils.py:22-130
(``DEFAULT_STREAM`` / ``get_stream_name`` / ``CUDAStreamPool`` / ``get_cuda_stream_pool``).

Stream 0 is always the stream the caller was on when it entered the compiled graph; indices 1.. are side streams taken from one
process-wide pool, so that several compiled graphs (forward and backward of every dynamo fragment) share the same few streams instead of
each creating its own."""
from __future__ import annotations

import functools
import threading

import torch

__all__ = ["DEFAULT_STREAM", "DEFAULT_STREAM_IDX", "ENTRANCE_EVENT", "EVENT_NAME_TEMPLATE", "STREAM_NAME_TEMPLATE", "CUDAStreamPool",
           "get_cuda_stream_pool", "get_stream_name"]

DEFAULT_STREAM: str = "default_stream"
DEFAULT_STREAM_IDX: int = 0
ENTRANCE_EVENT: str = "event0"
EVENT_NAME_TEMPLATE: str = "event{event_idx:d}"
STREAM_NAME_TEMPLATE: str = "stream{stream_idx:d}"


@functools.lru_cache(maxsize=None)
def get_stream_name(stream_idx: int) -> str:
    if stream_idx < 0:
        raise ValueError(f"stream index must be non-negative, got {stream_idx}")
    return DEFAULT_STREAM if stream_idx == DEFAULT_STREAM_IDX else STREAM_NAME_TEMPLATE.format(stream_idx=stream_idx)


class CUDAStreamPool:
    """``pool_size`` reusable side streams of one device. ``acquire`` / ``release`` hand streams out LIFO; as a context manager the
    pool makes one of its streams current for the ``with`` body. ``side_stream(i)`` is the stable mapping the generated programs use:
    side stream ``i`` (1-based) of every program is the same CUDA stream, created on first use."""

    def __init__(self, device: int | None = None, pool_size: int = 8) -> None:
        self.device, self.pool_size = device, pool_size
        self._lock = threading.Lock()
        self._all: list = []      # created lazily: constructing the pool must not touch CUDA
        self._free: list = []
        self._entered: list = []

    def _grow(self):
        if len(self._all) >= self.pool_size:
            raise RuntimeError(f"all {self.pool_size} streams of the pool are in use")
        s = torch.cuda.Stream(device=self.device)
        self._all.append(s)
        return s

    def acquire(self):
        with self._lock:
            return self._free.pop() if self._free else self._grow()

    def release(self, stream) -> None:
        if stream is not None:
            with self._lock:
                self._free.append(stream)

    def side_stream(self, stream_idx: int):
        """Stream behind ``stream<stream_idx>`` (1-based) of the generated programs."""
        if not 1 <= stream_idx <= self.pool_size:
            raise IndexError(f"side stream {stream_idx} outside a pool of {self.pool_size}")
        with self._lock:
            while len(self._all) < stream_idx:
                self._grow()
            return self._all[stream_idx - 1]

    def __enter__(self):
        s = self.acquire()
        ctx = torch.cuda.stream(s)
        ctx.__enter__()
        self._entered.append((s, ctx))
        return s

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        s, ctx = self._entered.pop()
        ctx.__exit__(exc_type, exc_val, exc_tb)
        self.release(s)


_pools: dict = {}


def get_cuda_stream_pool(device: int | None = None, pool_size: int = 32) -> CUDAStreamPool:
    """The process-wide pool of ``device`` (one per device; the reference keeps a single global one, _utils.py:111-130)."""
    pool = _pools.get(device)
    if pool is None:
        pool = _pools[device] = CUDAStreamPool(device=device, pool_size=pool_size)
    return pool
