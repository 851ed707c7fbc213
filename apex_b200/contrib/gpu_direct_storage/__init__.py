"""``gpu_direct_storage`` (reference apex/contrib/gpu_direct_storage/__init__.py, gds.cpp:45-165: cuFile read/write of a tensor's storage;
deprecated upstream). cuFile is not part of this image, so the same ``GDSFile.save_data / load_data`` API runs through a pinned bounce
buffer (the reference's own non-GDS fallback path)."""
from __future__ import annotations

import torch


class GDSFile:
    def __init__(self, filename: str, mode: str):
        assert mode in ("r", "w", "rw"), "mode must be one of r, w, rw"
        self.filename, self.mode = filename, mode
        self._f = None

    def __enter__(self):
        self._f = open(self.filename, {"r": "rb", "w": "wb", "rw": "r+b"}[self.mode])
        return self

    def __exit__(self, *a):
        self._f.close()
        self._f = None

    def save_data(self, t: torch.Tensor):
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=t.is_cuda)
        host.copy_(t)
        self._f.write(host.contiguous().view(torch.uint8).numpy().tobytes())

    def load_data(self, t: torch.Tensor):
        nbytes = t.numel() * t.element_size()
        buf = bytearray(self._f.read(nbytes))
        host = torch.frombuffer(buf, dtype=torch.uint8).view(t.dtype).view(t.shape)
        t.copy_(host)

    def save_data_no_gds(self, t):
        self.save_data(t)

    def load_data_no_gds(self, t):
        self.load_data(t)
