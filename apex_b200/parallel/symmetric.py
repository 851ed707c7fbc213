"""Symmetric memory over NVLink 5 / NVSwitch for in-kernel collectives.

``SymmetricMemory(nbytes, group)`` allocates the same-sized physical buffer on every rank of a node-local process group,
maps every peer's buffer into this process (P2P load/store from kernels) and, when the fabric supports it, binds all of
them to one NVSwitch multicast address (``multimem.ld_reduce`` / ``multimem.st``: in-switch reduction and broadcast).

Bootstrap only uses ``torch.distributed`` object collectives + POSIX fd passing; after construction no NCCL call is
needed to move data. This is the B200 replacement for the reference's three mechanisms: ``nccl_allocator`` (ncclMemAlloc
user buffers, apex/contrib/nccl_allocator/nccl_allocator.py:18-82), ``PeerMemoryPool`` (cudaIpc blob + all_gather of
handles, apex/contrib/peer_memory/peer_memory.py:6-115) and groupbn's IPC buffers (apex/contrib/groupbn/batch_norm.py:346-403).

Modes (env ``APEX_B200_SYMM``): ``vmm`` (default: cuMem VMM + fd exchange, multicast if available), ``ipc`` (legacy cudaIpc,
no multicast).
"""
from __future__ import annotations

import array
import ctypes
import os
import socket
import tempfile

import torch
import torch.distributed as dist

from .. import _lib

_lib.declare("ab_symm_granularity", "i i p")
_lib.declare("ab_symm_alloc", "i l l p p p")
_lib.declare("ab_symm_import", "i i l l p p")
_lib.declare("ab_symm_free", "l l l")
_lib.declare("ab_mc_create", "i l p p")
_lib.declare("ab_mc_import", "i p")
_lib.declare("ab_mc_add_device", "l i")
_lib.declare("ab_mc_bind_map", "l l i l l p")
_lib.declare("ab_ipc_alloc", "i l p p")
_lib.declare("ab_ipc_open", "i p p")
_lib.declare("ab_symm_barrier", "p i i i i p")

_SYS_pidfd_open = 434
_SYS_pidfd_getfd = 438


class _RawCudaMemory:
    """Minimal __cuda_array_interface__ carrier so torch can view driver-allocated memory without owning it."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self._owner = owner


def _tensor_from_ptr(ptr: int, nbytes: int, device, owner) -> torch.Tensor:
    return torch.as_tensor(_RawCudaMemory(ptr, nbytes, owner), device=device)


def _all_gather_obj(obj, group):
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def _dup_fd_from(pid: int, fd: int) -> int:
    libc = ctypes.CDLL(None, use_errno=True)
    pidfd = libc.syscall(_SYS_pidfd_open, pid, 0)
    if pidfd < 0:
        raise OSError(ctypes.get_errno(), "pidfd_open failed")
    try:
        new = libc.syscall(_SYS_pidfd_getfd, pidfd, fd, 0)
        if new < 0:
            raise OSError(ctypes.get_errno(), "pidfd_getfd failed")
        return new
    finally:
        os.close(pidfd)


def _exchange_fds(my_fd: int, group, tag: str) -> list[int]:
    """Every rank contributes one fd; returns this process's duplicates of every rank's fd (own slot = my_fd)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    infos = _all_gather_obj((os.getpid(), my_fd), group)
    ok = True
    fds = [-1] * world
    try:
        for r, (pid, fd) in enumerate(infos):
            fds[r] = my_fd if r == rank else _dup_fd_from(pid, fd)
    except OSError:
        ok = False
    oks = _all_gather_obj(ok, group)
    if all(oks):
        dist.barrier(group=group)  # nobody may close the source fd before every peer has duplicated it
        return fds
    for f in fds:
        if f >= 0 and f != my_fd:
            os.close(f)
    # fallback: SCM_RIGHTS over unix sockets
    d = tempfile.gettempdir()
    uid = _all_gather_obj(f"{os.getpid()}", group)[0]
    path = os.path.join(d, f"apexb200_{uid}_{tag}_{rank}.sock")
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    if os.path.exists(path):
        os.unlink(path)
    srv.bind(path)
    srv.listen(world + 1)
    dist.barrier(group=group)
    for r in range(world):
        if r == rank:
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(os.path.join(d, f"apexb200_{uid}_{tag}_{r}.sock"))
        c.sendmsg([rank.to_bytes(4, "little")], [(socket.SOL_SOCKET, socket.SCM_RIGHTS, array.array("i", [my_fd]))])
        c.close()
    fds = [-1] * world
    fds[rank] = my_fd
    for _ in range(world - 1):
        conn, _a = srv.accept()
        msg, anc, _f, _ad = conn.recvmsg(4, socket.CMSG_LEN(4))
        src = int.from_bytes(msg, "little")
        got = array.array("i")
        for level, typ, data in anc:
            if level == socket.SOL_SOCKET and typ == socket.SCM_RIGHTS:
                got.frombytes(data[:4])
        fds[src] = got[0]
        conn.close()
    dist.barrier(group=group)
    srv.close()
    os.unlink(path)
    return fds


def node_local(group=None) -> bool:
    names = _all_gather_obj(socket.gethostname(), group)
    return len(set(names)) == 1


class SymmetricMemory:
    """Same-size buffer on every rank, peer-mapped, optionally multicast-bound. ``world == 1`` degenerates to plain memory."""

    def __init__(self, nbytes: int, group=None, device=None, multicast: bool = True, zero: bool = True, tag: str = "m"):
        if not _lib.available():
            raise _lib.gpu_required_error("SymmetricMemory")
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.mode = os.environ.get("APEX_B200_SYMM", "vmm")
        self.mc_ptr = 0
        self._handles = []
        if self.world > 8:
            raise ValueError("SymmetricMemory spans one NVSwitch domain (<= 8 ranks)")
        want_mc = multicast and self.world > 1 and self.mode == "vmm" and os.environ.get("APEX_B200_NVLS", "1") == "1"
        if self.mode == "vmm":
            try:
                self._init_vmm(nbytes, dev, want_mc, tag)
            except Exception as e:  # noqa: BLE001
                oks = [False]
                self._vmm_error = e
            else:
                oks = [True]
            if self.world > 1:
                oks = _all_gather_obj(oks[0], group)
            if not all(oks):
                self.mode = "ipc"
        if self.mode == "ipc":
            self._init_ipc(nbytes, dev)
        self.local_ptr = self.peer_ptrs[self.rank]
        self.buffer = _tensor_from_ptr(self.local_ptr, self.nbytes, self.device, self)
        if zero:
            self.buffer.zero_()
            torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=group)

    # -- VMM + multicast ---------------------------------------------------------------------------------------------
    def _init_vmm(self, nbytes, dev, want_mc, tag):
        g = ctypes.c_uint64(0)
        _lib.fn("ab_symm_granularity")(dev, self.world if want_mc else 1, ctypes.addressof(g))
        gran = int(g.value)
        self.nbytes = (max(nbytes, 1) + gran - 1) // gran * gran
        self._gran = gran
        h, p, fd = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_int(-1)
        _lib.fn("ab_symm_alloc")(dev, self.nbytes, gran, ctypes.addressof(h), ctypes.addressof(p), ctypes.addressof(fd))
        self._handles.append((int(h.value), int(p.value)))
        self.peer_ptrs = [0] * self.world
        self.peer_ptrs[self.rank] = int(p.value)
        if self.world == 1:
            os.close(fd.value)
            return
        fds = _exchange_fds(fd.value, self.group, tag + "a")
        for r, f in enumerate(fds):
            if r == self.rank:
                continue
            hh, pp = ctypes.c_uint64(0), ctypes.c_uint64(0)
            _lib.fn("ab_symm_import")(dev, f, self.nbytes, gran, ctypes.addressof(hh), ctypes.addressof(pp))
            self._handles.append((int(hh.value), int(pp.value)))
            self.peer_ptrs[r] = int(pp.value)
            os.close(f)
        dist.barrier(group=self.group)
        os.close(fd.value)
        if want_mc:
            try:
                self._init_multicast(dev, tag)
            except Exception as e:  # noqa: BLE001
                self._mc_error = e
                self.mc_ptr = 0
            oks = _all_gather_obj(self.mc_ptr != 0, self.group)
            if not all(oks):
                self.mc_ptr = 0

    def _init_multicast(self, dev, tag):
        mch, mfd = ctypes.c_uint64(0), ctypes.c_int(-1)
        err = None
        if self.rank == 0:
            try:
                _lib.fn("ab_mc_create")(self.world, self.nbytes, ctypes.addressof(mch), ctypes.addressof(mfd))
            except Exception as e:  # noqa: BLE001
                err = e
        created = _all_gather_obj(err is None, self.group)[0]
        if not created:
            raise RuntimeError(f"multicast object creation failed: {err}")
        fds = _exchange_fds(mfd.value if self.rank == 0 else os.open(os.devnull, os.O_RDONLY), self.group, tag + "m")
        if self.rank != 0:
            _lib.fn("ab_mc_import")(fds[0], ctypes.addressof(mch))
        for r, f in enumerate(fds):
            if f >= 0:
                try:
                    os.close(f)
                except OSError:
                    pass
        _lib.fn("ab_mc_add_device")(int(mch.value), dev)
        dist.barrier(group=self.group)
        mp = ctypes.c_uint64(0)
        _lib.fn("ab_mc_bind_map")(int(mch.value), self._handles[0][0], dev, self.nbytes, self._gran, ctypes.addressof(mp))
        dist.barrier(group=self.group)
        self.mc_ptr = int(mp.value)
        self._mc_handle = int(mch.value)

    # -- legacy cudaIpc -------------------------------------------------------------------------------------------------
    def _init_ipc(self, nbytes, dev):
        self.nbytes = (max(nbytes, 1) + (1 << 21) - 1) // (1 << 21) * (1 << 21)
        p = ctypes.c_uint64(0)
        hbuf = ctypes.create_string_buffer(64)
        _lib.fn("ab_ipc_alloc")(dev, self.nbytes, ctypes.addressof(p), ctypes.addressof(hbuf))
        self.peer_ptrs = [0] * self.world
        self.peer_ptrs[self.rank] = int(p.value)
        if self.world == 1:
            return
        handles = _all_gather_obj(bytes(hbuf.raw), self.group)
        for r, hb in enumerate(handles):
            if r == self.rank:
                continue
            pp = ctypes.c_uint64(0)
            cb = ctypes.create_string_buffer(hb, 64)
            _lib.fn("ab_ipc_open")(dev, ctypes.addressof(cb), ctypes.addressof(pp))
            self.peer_ptrs[r] = int(pp.value)

    # -- views ----------------------------------------------------------------------------------------------------------
    def view(self, dtype: torch.dtype, numel: int, offset_bytes: int = 0) -> torch.Tensor:
        nb = numel * torch.empty((), dtype=dtype).element_size()
        assert offset_bytes + nb <= self.nbytes
        return self.buffer[offset_bytes:offset_bytes + nb].view(dtype)

    def peer_ptr_array(self, offset_bytes: int = 0):
        """ctypes uint64[world] of every rank's pointer (+offset) — pass with ctypes.addressof to kernels."""
        arr = (ctypes.c_uint64 * 8)(*[0] * 8)
        for r in range(self.world):
            arr[r] = self.peer_ptrs[r] + offset_bytes
        return arr

    @property
    def has_multicast(self) -> bool:
        return self.mc_ptr != 0


class SignalPad:
    """Per-group signal pad (epoch counters + tiny scratch) living in symmetric memory; one per process group."""

    _pads: dict = {}

    def __init__(self, group=None, device=None):
        words = 64 * 8 + 1024
        self.mem = SymmetricMemory(words * 4, group=group, device=device, multicast=False, zero=True, tag="p")
        self.ptrs = self.mem.peer_ptr_array()
        self.rank, self.world = self.mem.rank, self.mem.world
        self.epoch = 0
        self.device = self.mem.device
        self.done_ctr = torch.zeros(8, dtype=torch.int32, device=self.device)
        # device-resident epochs, one per channel: kernels that take an ``epoch_ctr`` pointer advance it themselves, so a captured
        # CUDA graph keeps counting on replay (a host-chosen epoch would be frozen into the graph)
        self.dev_epochs = torch.zeros(64, dtype=torch.int32, device=self.device)

    @classmethod
    def get(cls, group=None, device=None) -> "SignalPad":
        key = (id(group) if group is not None else 0, str(device))
        pad = cls._pads.get(key)
        if pad is None:
            pad = cls._pads[key] = SignalPad(group, device)
        return pad

    def next_epoch(self) -> int:
        self.epoch += 1
        return self.epoch

    def barrier(self, channel: int = 63):
        """Device-side barrier across the group on the current stream (no NCCL)."""
        if self.world == 1:
            return
        _lib.fn("ab_symm_barrier")(ctypes.addressof(self.ptrs), self.rank, self.world, self.next_epoch(), channel,
                                   _lib.stream_ptr(self.device))
