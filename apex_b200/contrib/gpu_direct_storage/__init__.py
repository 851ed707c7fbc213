"""``gpu_direct_storage`` (reference apex/contrib/gpu_direct_storage/__init__.py over gds.cpp:45-165: cuFile read / write of a tensor's
storage; deprecated upstream). cuFile is not part of this image, so ``GDSFile.save_data / load_data`` run on the native bounce pipeline
of ``csrc/file_io.cpp``: two pinned 32 MB staging buffers, the device copy of one chunk overlapped with the file I/O of the other;
host tensors are written / read in place. Within one ``with GDSFile(...)`` block successive calls append / continue (a single call
behaves like the reference: offset 0). Without the native library (never on a GPU machine) a plain-Python path is used."""
from __future__ import annotations

import ctypes
import os

import torch

from ... import _lib

_fns = None


def _native():
    """(write, read) entry points of the native runtime, or None when the library could not be loaded."""
    global _fns
    if _fns is None:
        try:
            w = _lib.raw_fn("ab_file_write", ctypes.c_longlong, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                                                               ctypes.c_int, ctypes.c_void_p])
            r = _lib.raw_fn("ab_file_read", ctypes.c_longlong, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                                                             ctypes.c_void_p])
            _fns = (w, r)
        except (RuntimeError, AttributeError):
            if torch.cuda.is_available():
                raise
            _fns = False
    return _fns or None


def _check(rc: int, what: str, path: str) -> int:
    if rc <= -100000:
        raise RuntimeError(f"gpu_direct_storage: {what} {path}: CUDA error {-rc - 100000}")
    if rc < 0:
        raise OSError(-rc, f"gpu_direct_storage: {what} failed: {os.strerror(-rc)}", path)
    return rc


class GDSFile:
    def __init__(self, filename: str, mode: str):
        assert mode in ("r", "w", "rw"), "mode must be one of r, w, rw"
        self.filename, self.mode = filename, mode
        self._pos = 0
        self._open = False

    def __enter__(self):
        self._pos, self._open = 0, True
        if self.mode == "w":
            open(self.filename, "wb").close()  # "w" truncates, as a cuFile handle opened with O_TRUNC
        return self

    def __exit__(self, *a):
        self._open = False

    def save_data(self, t: torch.Tensor):
        assert self._open and self.mode in ("w", "rw"), "file is not open for writing"
        src = t if t.is_contiguous() else t.contiguous()
        nbytes = src.numel() * src.element_size()
        fns = _native()
        if fns is None:
            with open(self.filename, "r+b") as f:
                f.seek(self._pos)
                f.write(src.cpu().view(torch.uint8).numpy().tobytes())
        else:
            stream = _lib.stream_ptr(src.device) if src.is_cuda else None
            _check(fns[0](self.filename.encode(), src.data_ptr(), nbytes, self._pos, int(src.is_cuda), 0, stream), "write", self.filename)
        self._pos += nbytes

    def load_data(self, t: torch.Tensor):
        assert self._open and self.mode in ("r", "rw"), "file is not open for reading"
        dst = t if t.is_contiguous() else torch.empty_like(t, memory_format=torch.contiguous_format)
        nbytes = dst.numel() * dst.element_size()
        fns = _native()
        if fns is None:
            with open(self.filename, "rb") as f:
                f.seek(self._pos)
                buf = bytearray(f.read(nbytes))
            got = len(buf)
            if got == nbytes:
                dst.copy_(torch.frombuffer(buf, dtype=torch.uint8).view(dst.dtype).view(dst.shape))
        else:
            stream = _lib.stream_ptr(dst.device) if dst.is_cuda else None
            got = _check(fns[1](self.filename.encode(), dst.data_ptr(), nbytes, self._pos, int(dst.is_cuda), stream), "read", self.filename)
        if got != nbytes:
            raise EOFError(f"gpu_direct_storage: {self.filename} holds {got} bytes at offset {self._pos}, tensor needs {nbytes}")
        if dst is not t:
            t.copy_(dst)
        self._pos += nbytes

    # the reference's explicit non-GDS variants: same pipeline here
    save_data_no_gds = save_data
    load_data_no_gds = load_data
