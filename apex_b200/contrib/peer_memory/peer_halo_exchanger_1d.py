"""1-D halo exchange through peer memory. Reference: apex/contrib/peer_memory/peer_halo_exchanger_1d.py:5-84 over
``push_pull_halos_1d`` (peer_memory_cuda.cu:146-295: 16-byte "flit" stores carrying payload + flag into the neighbour's buffer and
a volatile spin on the local one, cooperative launch). Here the whole exchange is ONE launch of csrc/halo_exchange.cu: the two
outgoing slabs are packed straight into the neighbours' transfer buffers with 16-byte P2P stores over NVLink, an epoch word per
neighbour (release / acquire at .sys scope) orders them, and the incoming slabs are unpacked into the halo rows; transfer buffers
are double-buffered by exchange parity, so there is no trailing barrier and no host synchronisation."""
from __future__ import annotations

import torch

import ctypes

from ... import _lib
from ...parallel.symmetric import SignalPad

_lib.declare("ab_halo_exchange_1d", "p i p p p p p p p i i i i i i i p l p i p")


class PeerHaloExchanger1d:
    def __init__(self, ranks, rank_in_group, peer_pool, half_halo):
        self.peer_group_size = len(ranks)
        self.ranks = ranks
        self.peer_rank = rank_in_group
        self.low_neighbor = (self.peer_rank + self.peer_group_size - 1) % self.peer_group_size
        self.high_neighbor = (self.peer_rank + 1) % self.peer_group_size
        self.low_zero = self.peer_rank == 0
        self.high_zero = self.peer_rank == self.peer_group_size - 1
        self.peer_pool = peer_pool
        self.half_halo = half_halo
        self.pad = SignalPad.get(peer_pool.group, peer_pool.mem.device)

    def _slices(self, y, H_split, explicit_nhwc):
        hh = self.half_halo
        dim = (1 if explicit_nhwc else 2) if H_split else (2 if explicit_nhwc else 3)
        L = y.shape[dim] - 2 * hh
        sl = lambda a, b: tuple(slice(a, b) if d == dim else slice(None) for d in range(4))
        return y[sl(hh, 2 * hh)], y[sl(0, hh)], y[sl(L, L + hh)], y[sl(L + hh, L + 2 * hh)]

    def __call__(self, y, H_split=True, explicit_nhwc=False, numSM=0, diagnostics=False):
        low_out, low_in, high_out, high_in = self._slices(y, H_split, explicit_nhwc)
        if y.is_cuda and _lib.available() and y.element_size() in (2, 4):
            return self._fused(y, low_out, low_in, high_out, high_in, numSM)
        # tx[r] is the buffer living on rank r: [0] receives from its low neighbour, [1] from its high neighbour
        tx = self.peer_pool.allocate_peer_tensors([2] + list(low_out.shape), y.dtype, False, True)
        if not self.low_zero:
            tx[self.low_neighbor][1].copy_(low_out)    # my low-side interior rows are the low neighbour's HIGH input halo
        if not self.high_zero:
            tx[self.high_neighbor][0].copy_(high_out)
        self.pad.barrier(channel=40)                    # every push has landed everywhere
        mine = tx[self.peer_rank]
        if self.low_zero:
            low_in.zero_()
        else:
            low_in.copy_(mine[0])
        if self.high_zero:
            high_in.zero_()
        else:
            high_in.copy_(mine[1])
        self.pad.barrier(channel=41)                    # buffers may be reused after this

    def _fused(self, y, low_out, low_in, high_out, high_in, numSM):
        nbytes = low_out.numel() * y.element_size()
        key = (nbytes, y.dtype)
        st = getattr(self, "_tx", None)
        if st is None or st[0] != key:
            # [parity][side][slab] in the STATIC part of the pool: lives across exchanges (double buffering by parity)
            tx = self.peer_pool.allocate_peer_tensors([2, 2, low_out.numel()], y.dtype, False, False)
            self._tx = st = (key, tx, torch.zeros(1, dtype=torch.int32, device=y.device), [0])
        _, tx, ticket, it = st
        # the epoch — and with it the parity of the double-buffered transfer slabs — lives in device memory (pad.dev_epochs[42]): the
        # kernel advances it, so the exchange can be captured in a CUDA graph and replayed
        par = 0
        parity_stride = (tx[0][1].data_ptr() - tx[0][0].data_ptr())
        # innermost-last ordering of the slab dims so that 16-byte vectors run along the unit-stride dim
        order = sorted(range(4), key=lambda d: (-low_out.stride(d), d))
        base = y.untyped_storage().data_ptr()
        es = y.element_size()
        off = lambda t: (t.data_ptr() - base) // es
        meta = (ctypes.c_longlong * 12)(*[low_out.shape[d] for d in order], *[low_out.stride(d) for d in order],
                                        off(low_out), off(low_in), off(high_out), off(high_in))
        pads = self.pad.mem.peer_ptrs
        me, lo, hi = self.peer_rank, self.low_neighbor, self.high_neighbor
        _lib.fn("ab_halo_exchange_1d")(base, es, ctypes.addressof(meta), tx[lo][par].data_ptr(), tx[hi][par].data_ptr(),
                                       tx[me][par].data_ptr(), pads[lo], pads[hi], pads[me], me, lo, hi, int(not self.low_zero),
                                       int(not self.high_zero), 42, 0, self.pad.dev_epochs[42:43].data_ptr(), parity_stride,
                                       ticket.data_ptr(), int(numSM),
                                       _lib.stream_ptr(y.device))
