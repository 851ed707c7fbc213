"""DistributedFusedAdam — ZeRO-2 Adam (optimizer state and gradients sharded over data-parallel ranks).

Public behaviour follows the reference ``apex.contrib.optimizers.DistributedFusedAdam``
(apex/contrib/optimizers/distributed_fused_adam.py:270-3488): Adam hyper-parameters per param group, ``dtype`` (state),
``grad_sync_dtype``, ``param_sync_dtype``, ``average_grad_sync``, ``bucket_cap_mb``, ``no_sync``, ``grad_sync``, ``param_sync``,
``grad_norm`` / ``clip_grad_norm`` (deferred through a device ``_grad_scale``), ``unscale_grads`` (GradScaler hook, overflow
detected through the gradient norm), ``grad_buffer_view``, ``zero_grad``, reshardable (world-size independent) ``state_dict``.

What is different is HOW a step runs on B200 (csrc/dist_adam.cu):
  * gradients and low-precision parameters live in contiguous buffers on a SYMMETRIC HEAP (parallel/symmetric.py) that all
    ranks of the node map; ``param.grad`` / ``param.data`` are views into them, so there is no copy-into-bucket pass and no
    bucket->param copy-out pass (reference :1600-1666, :1716-1767);
  * ``step()`` is ONE kernel per parameter segment: it pulls this rank's gradient shard from every peer over NVLink (or one
    ``multimem.ld_reduce`` through NVSwitch), scales, accumulates the gradient norm, applies Adam to the fp32 shard and pushes
    the new parameters into every rank's parameter buffer (or one ``multimem.st``) — no NCCL call, no intermediate buffers;
  * when the update needs the global norm first (clipping, GradScaler) the same kernel runs as two phases (RS+norm, Adam+AG).
The NCCL / gloo implementation of the same algorithm (``fused_collectives=False``; always used on CPU, across nodes, for
exotic dtypes) is kept as the in-repo baseline and as the oracle for the fused path.
"""
from __future__ import annotations

import contextlib
import ctypes
from dataclasses import dataclass
from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist

from ... import _lib
from ...ops import amp_C
from ...ops import reference as ref

_lib.declare("ab_dist_adam_step", "i i p p p l l p p p p p l i i i i i i i p i i i p p p p f f f f f i i i f p p p p p i i p i i i i p")
_lib.declare("ab_symm_gate", "p i i p i p")

_CHUNK = 2048  # elements handled by one CTA work item in csrc/dist_adam.cu
_ALIGN = 64    # every parameter starts on a 64-element boundary of the flat space


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _pack_msb(src: torch.Tensor, dst: torch.Tensor) -> None:
    """Bit transport between a floating-point tensor and an INTEGER ``param_sync_dtype`` (reference :2810-2824): the most significant bytes of
    every element are copied (little-endian), missing low bytes are zero. int32 / int64 carry fp32 losslessly, uint8 keeps sign + 7 exponent bits."""
    a = src.contiguous().unsqueeze(-1).view(torch.uint8)
    b = dst.unsqueeze(-1).view(torch.uint8)
    n = min(a.size(-1), b.size(-1))
    if n < b.size(-1):
        b[..., :-n].zero_()
    b[..., -n:].copy_(a[..., -n:])


class _Segment:
    """All parameters of one (param group, dtype triple): a flat space cut into buckets, each bucket sharded D ways."""

    def __init__(self, opt, group_idx, params, dtype, grad_dtype, param_dtype):
        self.opt, self.group_idx, self.params = opt, group_idx, params
        self.dtype, self.grad_dtype, self.param_dtype = dtype, grad_dtype, param_dtype
        D, rank = opt.distributed_size, opt.distributed_rank
        self.D, self.rank = D, rank
        # world size 1: every parameter starts on a chunk boundary so that the step kernel can read a parameter's gradient in place
        # (one source per chunk, see ``direct``); sharded layouts pack at 64 elements
        align = _CHUNK if D == 1 else _ALIGN
        off, self.offsets = 0, []
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + align - 1) // align * align
        self.numel = off
        gran = D * _CHUNK
        esize = max(torch.empty((), dtype=grad_dtype).element_size(), torch.empty((), dtype=param_dtype).element_size())
        cap = max(gran, int(opt.bucket_cap_mb * 1024 * 1024 / esize) // gran * gran)
        self.bucket_elems = min(cap, (max(off, 1) + gran - 1) // gran * gran)
        self.n_buckets = (max(off, 1) + self.bucket_elems - 1) // self.bucket_elems
        self.padded = self.n_buckets * self.bucket_elems
        self.shard_elems = self.bucket_elems // D
        self.local_elems = self.n_buckets * self.shard_elems
        dev = opt.device
        self.fused = opt._fused_ok(dtype, grad_dtype, param_dtype)
        self.symm_g = self.symm_p = None
        gsz = torch.empty((), dtype=grad_dtype).element_size()
        psz = torch.empty((), dtype=param_dtype).element_size()
        if self.fused and D > 1:
            from ...parallel.symmetric import SymmetricMemory

            self.symm_g = SymmetricMemory(self.padded * gsz, group=opt.distributed_process_group, device=dev, tag=f"g{group_idx}")
            self.symm_p = SymmetricMemory(self.padded * psz, group=opt.distributed_process_group, device=dev, tag=f"w{group_idx}")
            self.grad_buf = self.symm_g.view(grad_dtype, self.padded)
            self.param_buf = self.symm_p.view(param_dtype, self.padded)
        else:
            self.grad_buf = torch.zeros(self.padded, dtype=grad_dtype, device=dev)
            self.param_buf = torch.zeros(self.padded, dtype=param_dtype, device=dev)
        # local optimizer state (bucket-major shard layout)
        self.exp_avg = torch.zeros(self.local_elems, dtype=dtype, device=dev)
        self.exp_avg_sq = torch.zeros(self.local_elems, dtype=dtype, device=dev)
        self.master = torch.zeros(self.local_elems, dtype=dtype, device=dev) if opt.store_params else None
        self.remainders = torch.zeros(self.local_elems, dtype=torch.int16, device=dev) if opt.store_param_remainders else None
        self.reduced = None
        # one (local, global) sum-of-squares pair per reduce-scatter launch: row b0 belongs to the launch that started at bucket b0
        self.norm_out = torch.zeros(self.n_buckets, 2, dtype=torch.float32, device=dev)
        self.norm_rows: list[int] = []   # rows written since the last zero_grad / step
        self.norm_partials = torch.zeros(1024, dtype=torch.float32, device=dev) if dev.type == "cuda" else None
        self.bucket_synced = [False] * self.n_buckets   # reduced shard holds this step's reduce-scattered grads of bucket b
        # overlap_grad_sync bookkeeping: parameters that intersect each bucket, and how many of them still owe a gradient this step
        self.param_buckets = []
        self.bucket_nparams = [0] * self.n_buckets
        for p, off0 in zip(params, self.offsets):
            bs = range(off0 // self.bucket_elems, (off0 + max(p.numel(), 1) - 1) // self.bucket_elems + 1)
            self.param_buckets.append(bs)
            for b in bs:
                self.bucket_nparams[b] += 1
        self.bucket_pending = list(self.bucket_nparams)
        self.bucket_stepped = [False] * self.n_buckets  # overlap_step_with_backward: the fused step of bucket b already ran this step
        self.written = [False] * len(params)   # zero_grad(set_to_none=True): the buffer region of parameter i holds this step's gradient
        # world size 1, zero_grad(set_to_none=True): gradients that autograd hands over are not even copied — the step kernel reads them
        # where they are (csrc/dist_adam.cu src_tab); ``live`` keeps them alive until the step has consumed them
        self.live = [None] * len(params)
        self.src_tab_host = None
        self.src_tab_dev = None
        self.cast_params = [p for p in params if (not param_dtype.is_floating_point) or p.dtype != param_dtype]
        # overlap_param_sync: per-(bucket, rank) ready flags in memory every rank can write (symmetric heap), local chunk counters
        self.region = None
        self.ready_ptrs = None
        if self.fused and getattr(opt, "overlap_param_sync", False) and dev.type == "cuda":
            from ...parallel import param_sync as _ps

            nb = self.n_buckets * 8 * 4
            if D > 1:
                from ...parallel.symmetric import SymmetricMemory

                self.symm_flags = SymmetricMemory(nb, group=opt.distributed_process_group, device=dev, multicast=False, tag=f"f{group_idx}")
                self.ready_ptrs = self.symm_flags.peer_ptr_array()
                flags_ptr = self.symm_flags.local_ptr
            else:
                self.flags = torch.zeros(self.n_buckets * 8, dtype=torch.int32, device=dev)
                self.ready_ptrs = (ctypes.c_uint64 * 8)(self.flags.data_ptr(), 0, 0, 0, 0, 0, 0, 0)
                flags_ptr = self.flags.data_ptr()
            self.bucket_ctr = torch.zeros(self.n_buckets, dtype=torch.int32, device=dev)
            self.region = _ps.register(_ps.Region(self.param_buf, flags_ptr, D, self.bucket_elems))
        self.scales = None
        if getattr(opt, "with_scaled_states", False):
            # per-(parameter x shard) fragment scale factors for the 16-bit state (reference :2693-2774,2833-2860): element i of the
            # local shard belongs to fragment frag_index[i]; the padding shares one extra fragment
            frags = self.fragments()
            idx = torch.full((self.local_elems,), len(frags), dtype=torch.int64, device=dev)
            for f, (_, s0, n) in enumerate(frags):
                idx[s0:s0 + n] = f
            self.frag_index = idx
            self.scales = {k: torch.ones(len(frags) + 1, dtype=torch.float32, device=dev) for k in ("param", "exp_avg", "exp_avg_sq")}
        self._init_views()

    def fragments(self):
        """(parameter index, start in the local shard arrays, length) of every (parameter x this rank's shard) intersection."""
        out = []
        B, Sb, r = self.bucket_elems, self.shard_elems, self.rank
        for pi, (p, off) in enumerate(zip(self.params, self.offsets)):
            lo, hi = off, off + p.numel()
            for b in range(lo // B, (hi - 1) // B + 1):
                s_lo, s_hi = b * B + r * Sb, b * B + (r + 1) * Sb
                a, z = max(lo, s_lo), min(hi, s_hi)
                if z > a:
                    out.append((pi, b * Sb + (a - s_lo), z - a))
        return out

    def load_scaled(self, key: str) -> torch.Tensor:
        """fp32 value of a scaled 16-bit state tensor."""
        t = {"param": self.master, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}[key]
        return t.float() * self.scales[key][self.frag_index]

    def store_scaled(self, key: str, value: torch.Tensor) -> None:
        """value (fp32) -> 16-bit state with a fresh per-fragment scale = absmax / largest finite value of the state dtype."""
        t = {"param": self.master, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}[key]
        amax = torch.zeros_like(self.scales[key]).scatter_reduce_(0, self.frag_index, value.abs(), "amax", include_self=True)
        # scale = absmax / largest finite value, floored at the smallest NORMAL fp32: for bf16 (fp32's exponent range) the quotient is
        # subnormal below absmax ~ 4 and underflows to zero below ~5e-7 (second moments get there), which would turn the division below
        # into inf / nan; with the floor the stored values still fit (|value| / scale <= max) and the scale keeps its full mantissa.
        # (The reference divides by max / 2 and zeroes the state when the scale underflows, :2834-2860.)
        sc = torch.where(amax > 0, (amax / torch.finfo(t.dtype).max).clamp_(min=torch.finfo(torch.float32).tiny), torch.ones_like(amax))
        self.scales[key].copy_(sc)
        t.copy_((value / sc[self.frag_index]).to(t.dtype))

    # flat <-> shard helpers ------------------------------------------------------------------------------------------
    def shard_view(self, full: torch.Tensor, r: Optional[int] = None) -> torch.Tensor:
        """[n_buckets, Sb] strided view of rank r's shard inside a full-size flat buffer."""
        r = self.rank if r is None else r
        return full.view(self.n_buckets, self.D, self.shard_elems)[:, r, :]

    def _init_views(self):
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                n = p.numel()
                pv = self.param_buf[off:off + n].view(p.shape)
                if self.param_dtype.is_floating_point:
                    pv.copy_(p.detach().to(self.param_dtype))
                else:
                    _pack_msb(p.detach(), pv)
                if p.dtype == self.param_dtype:
                    p.data = pv  # the model weight IS the all-gather destination
                self.opt._param_view[id(p)] = pv
                self.opt._grad_view[id(p)] = self.grad_buf[off:off + n].view(p.shape)
            has_init = any(id(p) in self.opt._init_values for p in self.params) or not self.param_dtype.is_floating_point
            if self.master is not None and not has_init:
                # master := current parameter values of this rank's shard (strided read, no full-size temporary)
                self.master.view(self.n_buckets, self.shard_elems).copy_(self.shard_view(self.param_buf))
            elif self.master is not None:
                # fp32 master initialised from user-provided higher precision values
                full = torch.zeros(self.padded, dtype=self.dtype, device=self.opt.device)
                for p, off in zip(self.params, self.offsets):
                    src = self.opt._init_values.get(id(p), p.detach())
                    full[off:off + p.numel()].copy_(src.reshape(-1).to(self.dtype))
                self.master.view(self.n_buckets, self.shard_elems).copy_(self.shard_view(full))
                del full
            elif self.remainders is not None:
                full = torch.zeros(self.padded, dtype=torch.float32, device=self.opt.device)
                for p, off in zip(self.params, self.offsets):
                    src = self.opt._init_values.get(id(p), p.detach())
                    full[off:off + p.numel()].copy_(src.reshape(-1).float())
                sh = self.shard_view(full).contiguous().view(-1)
                bits = sh.view(torch.int32)
                lo = (bits & 0xFFFF).to(torch.int32)
                lo = torch.where(lo >= 32768, lo - 65536, lo).to(torch.int16)
                self.remainders.copy_(lo)
                del full

    @property
    def synced(self) -> bool:
        return all(self.bucket_synced)

    @synced.setter
    def synced(self, v: bool):
        self.bucket_synced = [bool(v)] * self.n_buckets
        if not v:
            self.norm_rows = []
            self.bucket_pending = list(self.bucket_nparams)
            self.bucket_stepped = [False] * self.n_buckets
            self.written = [False] * len(self.params)
            self.live = [None] * len(self.params)

    def direct_ok(self, g: torch.Tensor) -> bool:
        return (self.D == 1 and self.fused and g.is_cuda and g.dtype == self.grad_dtype and g.is_contiguous() and g.numel() % 8 == 0
                and g.data_ptr() % 16 == 0 and not self.opt.overlap_step_with_backward)

    def source_table(self):
        """Device table for this step's launch, or (None, 0) when every gradient sits in the contiguous buffer."""
        if not any(t is not None for t in self.live):
            return None, 0
        n = len(self.params)
        if self.src_tab_host is None:
            self.src_tab_host = torch.empty(n, 3, dtype=torch.int64).pin_memory()
            self.src_tab_dev = torch.empty(n, 3, dtype=torch.int64, device=self.opt.device)
        rows = []
        esz = self.grad_buf.element_size()
        base = self.grad_buf.data_ptr()
        for i, p in enumerate(self.params):
            t = self.live[i]
            if t is not None:
                rows.append((self.offsets[i] // _CHUNK, t.data_ptr(), p.numel()))
            else:   # buffer slot: padded (and zero beyond the parameter) up to the next chunk boundary, so whole vectors are valid
                rows.append((self.offsets[i] // _CHUNK, base + self.offsets[i] * esz, (p.numel() + _CHUNK - 1) // _CHUNK * _CHUNK))
        self.src_tab_host.copy_(torch.tensor(rows, dtype=torch.int64))
        self.src_tab_dev.copy_(self.src_tab_host, non_blocking=True)
        return self.src_tab_dev.data_ptr(), n

    def grad_sq(self) -> torch.Tensor:
        """Global sum of squares of the gradients reduced so far this step (device scalar)."""
        if len(self.norm_rows) == 1:
            return self.norm_out[self.norm_rows[0], 1]
        return self.norm_out[self.norm_rows, 1].sum()

    def attach_grads(self):
        for p in self.params:
            p.grad = self.opt._grad_view[id(p)] if p.dtype == self.grad_dtype else None


class DistributedFusedAdam(torch.optim.Optimizer):
    """ZeRO-2 Adam / AdamW: optimizer state and gradient reduction sharded over the data-parallel group, parameters all-gathered
    after the update. Reference: apex/contrib/optimizers/distributed_fused_adam.py:477-3488 (same constructor arguments, methods
    ``init_params / zero_grad / grad_buffer_view / no_sync / grad_sync / param_sync / grad_norm / clip_grad_norm / unscale_grads /
    step(grad_scaler=) / state_dict / load_state_dict``).

    On one NVSwitch node (<= 8 ranks) with fp32 state and 16/32-bit float gradients and parameters, ``step()`` is ONE kernel
    (csrc/dist_adam.cu): reduce-scatter by P2P pulls or ``multimem.ld_reduce``, gradient norm, Adam on the fp32 shard, parameter
    all-gather by P2P pushes or ``multimem.st`` — gradients and parameters live in symmetric-heap buffers that every rank maps.
    Everything else (CPU / gloo, >8 ranks, redundant groups, 16-bit or scaled state, parameter remainders) takes the bucketed
    NCCL / gloo path built on the multi-tensor kernels; both paths produce the same numbers.

    The fused step spins on its peers: do not keep an NCCL collective of the same process in flight on another stream while
    ``step()`` runs — a ZeRO training loop does not, its only collectives are inside the step."""

    def __init__(self, params, lr: float = 1e-3, bias_correction: bool = True, betas=(0.9, 0.999), eps: float = 1e-8,
                 adam_w_mode: bool = True, weight_decay: float = 0.0, amsgrad: bool = False, dtype: torch.dtype = torch.float32,
                 grad_sync_dtype: Optional[torch.dtype] = None, param_sync_dtype: Optional[torch.dtype] = None, device="cuda",
                 process_group=None, distributed_process_group=None, redundant_process_group=None, average_grad_sync: bool = True,
                 overlap_grad_sync: bool = True, overlap_param_sync: bool = False, bucket_cap_mb: float = 100.0,
                 pipeline_size: int = 2, contiguous_param_buffer: bool = True, contiguous_grad_buffer: bool = True,
                 store_params: bool = True, store_param_remainders: bool = False, with_scaled_states: bool = False,
                 nccl_ub: bool = False, capturable: bool = False, fused_collectives="auto", overlap_step_with_backward: bool = False):
        if amsgrad:
            raise RuntimeError("DistributedFusedAdam does not support the AMSGrad variant.")
        if with_scaled_states and (dtype not in (torch.float16, torch.bfloat16) or not store_params or store_param_remainders):
            raise RuntimeError("with_scaled_states needs 16-bit optimizer state (dtype=fp16/bf16) and store_params=True")
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.adam_w_mode = adam_w_mode
        self.dtype, self._grad_sync_dtype, self._param_sync_dtype = dtype, grad_sync_dtype, param_sync_dtype
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.process_group = process_group
        self.distributed_process_group = distributed_process_group if distributed_process_group is not None else process_group
        self.redundant_process_group = redundant_process_group
        self.distributed_size, self.distributed_rank = _world(self.distributed_process_group)
        self.redundant_size = _world(redundant_process_group)[0] if redundant_process_group is not None else 1
        self.average_grad_sync = average_grad_sync
        self.overlap_grad_sync, self.overlap_param_sync = overlap_grad_sync, overlap_param_sync
        self.bucket_cap_mb, self.pipeline_size = bucket_cap_mb, pipeline_size
        self.contiguous_param_buffer, self.contiguous_grad_buffer = True, True  # always contiguous in this implementation
        if store_param_remainders:
            store_params = False
            if dtype != torch.float32:
                raise RuntimeError("store_param_remainders requires fp32 optimizer state")
        self.store_params, self.store_param_remainders = store_params, store_param_remainders
        self.capturable = capturable
        self.with_scaled_states = with_scaled_states
        self.nccl_ub = nccl_ub
        self._fused_request = fused_collectives
        # B200 extension (not in the reference): run the WHOLE step of a bucket (reduce-scatter + Adam + parameter push, one
        # kernel) from the post-accumulate-grad hook as soon as backward has produced the bucket's gradients, on a side stream, so the
        # HBM- / link-bound optimizer hides under the tensor-core-bound backward; step() then only joins. The arithmetic is the same
        # kernel on the same data. Needs the global gradient norm to be irrelevant to the update: no clip_grad_norm, no GradScaler.
        self.overlap_step_with_backward = bool(overlap_step_with_backward)
        self._step_bumped = False
        self._param_view, self._grad_view, self._init_values = {}, {}, {}
        self._segments: list[_Segment] = []
        self._grad_scale = torch.ones([], dtype=torch.float32, device=self.device)
        self._dtype_overrides: dict = {}
        self._grad_norm = None
        self._dummy_overflow_buf = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._pad = None
        self._step_supports_amp_scaling = True
        self._inited = False
        self._sync_enabled = True
        self._side_stream = None          # overlap_grad_sync: per-bucket reduce-scatter launches run here while backward continues
        self._hook_handles = []
        self._overlap_launched = False
        self._push_in_flight = False
        self._overlap_ok = False
        self._steal = False    # zero_grad(set_to_none=True) is in effect: gradients arrive as fresh tensors and are copied into the buffer
        self.last_nvls = False
        self.kernel_launches = 0  # number of csrc/dist_adam.cu launches so far (bench.py reports it)
        self._last_grad_norm = None
        self._last_norm_rows: dict = {}
        if capturable:
            # graph-capturable: learning rate and step count are device tensors read by the kernel (reference :576-582)
            for group in self.param_groups:
                group["lr"] = torch.as_tensor(group["lr"], dtype=torch.float32, device=self.device).reshape(1).clone()
                group["step"] = torch.zeros(1, dtype=torch.int32, device=self.device)
        if self.device.type == "cuda" and not _lib.available():
            raise _lib.gpu_required_error("DistributedFusedAdam")
        self._decide_fused()
        # the reference broadcasts parameters from rank 0 at construction (:846-862)
        if self.distributed_size > 1 or self.redundant_size > 1:
            for group in self.param_groups:
                for p in group["params"]:
                    dist.broadcast(p.data, src=dist.get_global_rank(self.process_group, 0) if self.process_group is not None else 0,
                                   group=self.process_group)

    # ---------------------------------------------------------------------------------------------------------------
    def _decide_fused(self):
        req = self._fused_request
        ok = self.device.type == "cuda" and self.redundant_size == 1 and self.distributed_size <= 8
        if ok and self.distributed_size > 1:
            from ...parallel.symmetric import node_local

            ok = node_local(self.distributed_process_group)
        if req is True and not ok:
            raise RuntimeError("fused_collectives=True needs CUDA, one node, <= 8 ranks and no redundant group")
        self.fused_collectives = ok if req in ("auto", True) else False

    def _fused_ok(self, dtype, grad_dtype, param_dtype) -> bool:
        """The one-kernel path keeps fp32 state and 16/32-bit float grads/params; everything else takes the generic path."""
        f = (torch.float32, torch.float16, torch.bfloat16)
        pairs = {(torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16), (torch.float32, torch.float32),
                 (torch.bfloat16, torch.float32), (torch.float16, torch.float32), (torch.float32, torch.bfloat16),
                 (torch.float32, torch.float16)}
        # fp32 master weights, or (store_param_remainders) bf16 parameters + int16 remainders reassembled inside the kernel
        master_ok = self.store_params or (self.store_param_remainders and param_dtype == torch.bfloat16)
        return (self.fused_collectives and not self.with_scaled_states and dtype == torch.float32 and master_ok and grad_dtype in f and param_dtype in f
                and (grad_dtype, param_dtype) in pairs)

    def init_params(self, params: Optional[Iterable[torch.nn.Parameter]] = None, dtype: Optional[torch.dtype] = None,
                    grad_sync_dtype: Optional[torch.dtype] = None, param_sync_dtype: Optional[torch.dtype] = None) -> None:
        """Lay out parameters, allocate buffers and state (lazily called by the first step / zero_grad). Called with ``params`` and dtypes
        BEFORE that point it records per-parameter overrides of the optimizer-wide state / gradient-sync / parameter-sync dtypes, as the
        reference allows (:1228-1273; e.g. a few fp32 parameters inside a bf16 model); the layout itself is still built in one go."""
        if self._inited:
            return
        if params is not None and (dtype or grad_sync_dtype or param_sync_dtype):
            for p in params:
                self._dtype_overrides[id(p)] = (dtype, grad_sync_dtype, param_sync_dtype)
            return
        for gi, group in enumerate(self.param_groups):
            keyed: dict = {}
            for p in group["params"]:
                if not p.requires_grad:
                    continue
                sd, gd, pd = self._dtype_overrides.get(id(p), (None, None, None))
                gd = gd or self._grad_sync_dtype or p.dtype
                pd = pd or self._param_sync_dtype or p.dtype
                keyed.setdefault((sd or self.dtype, gd, pd), []).append(p)
            for (sd, gd, pd), ps in keyed.items():
                self._segments.append(_Segment(self, gi, ps, sd, gd, pd))
        padded = sum(seg.padded for seg in self._segments)
        # same check and wording as the reference (:1540-1549); a single bucket is already shrunk to its data, only its granule padding is left
        if padded and any(seg.n_buckets > 1 for seg in self._segments) and sum(seg.numel for seg in self._segments) / padded < 0.7:
            import warnings

            warnings.warn(f"Only {sum(seg.numel for seg in self._segments) / padded:.1%} of buckets are used. "
                          "Consider decreasing the bucket_cap_mb argument.")
        if self.fused_collectives and self.distributed_size > 1:
            from ...parallel.symmetric import SignalPad

            self._pad = SignalPad.get(self.distributed_process_group, self.device)
        self._done_ctr = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._done_ctr_side = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._inited = True
        self._install_overlap_hooks()
        self._collect_grads()  # gradients that already exist (assigned before the first step) are folded in, not dropped
        for seg in self._segments:
            seg.attach_grads()

    # ---- overlap_grad_sync: reduce-scatter every bucket as soon as backward has produced all of its gradients -------------
    def _install_overlap_hooks(self):
        """Reference :899-913, :1827-1875: post-accumulate-grad hooks start a bucket's gradient synchronisation on a side stream
        while backward continues. Here the synchronisation of a bucket is one MODE_RS launch of csrc/dist_adam.cu over that bucket
        (in-kernel pulls / multimem.ld_reduce, result in the fp32 reduced shard); step() then only runs Adam + the parameter push
        for everything that was reduced during backward. Every rank must produce gradients in the same order (same model)."""
        if not hasattr(torch.Tensor, "register_post_accumulate_grad_hook"):
            return
        self._overlap_ok = bool(self.overlap_grad_sync and self.device.type == "cuda" and self.distributed_size > 1)
        for si, seg in enumerate(self._segments):
            for pi, p in enumerate(seg.params):
                self._hook_handles.append(p.register_post_accumulate_grad_hook(self._make_hook(si, pi)))

    def _make_hook(self, si: int, pi: int):
        def hook(param):
            seg = self._segments[si]
            g, gv = param.grad, self._grad_view[id(param)]
            if g is None:
                return
            if g is not gv and g.data_ptr() != gv.data_ptr():
                if not self._steal:
                    return   # a user-assigned gradient tensor: step() folds it in and syncs then
                # zero_grad(set_to_none=True): autograd handed over a fresh gradient tensor instead of accumulating into the (zeroed)
                # buffer view; ONE copy moves it into the buffer (4 B/element instead of zero + read-modify-write = 8) and frees it
                if seg.live[pi] is not None:
                    seg.live[pi].add_(g)                 # later micro-batch: accumulate into the tensor the kernel will read
                elif seg.written[pi]:
                    gv.add_(g)
                elif seg.direct_ok(g):
                    seg.live[pi] = g                     # world size 1: no copy at all, the step kernel reads it in place
                    seg.written[pi] = True
                else:
                    gv.copy_(g)
                    seg.written[pi] = True
                param.grad = None
            step_now = self.overlap_step_with_backward and seg.fused and self.device.type == "cuda"
            if not (self._sync_enabled and seg.fused and (self._overlap_ok or step_now)):
                return
            for b in seg.param_buckets[pi]:
                seg.bucket_pending[b] -= 1
                if seg.bucket_pending[b] == 0 and not seg.bucket_synced[b] and not seg.bucket_stepped[b]:
                    self._overlap_sync_bucket(seg, b, 0 if step_now else 1)
        return hook

    def _bump_step(self):
        """Advance the step counters once per optimizer step (bias correction of every bucket launched for this step uses it)."""
        if self._step_bumped:
            return
        self._step_bumped = True
        for group in self.param_groups:
            if self.capturable:
                group["step"] += (self._dummy_overflow_buf != 1).to(torch.int32)
            else:
                group["step"] = group.get("step", 0) + 1

    def _overlap_sync_bucket(self, seg, b: int, mode: int = 1):
        import os as _os

        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        group = self.param_groups[seg.group_idx]
        if mode == 0:
            if not self._step_bumped:
                self._dummy_overflow_buf.zero_()
            self._bump_step()
        cur = torch.cuda.current_stream(self.device)
        self._side_stream.wait_stream(cur)          # the gradients of this bucket are complete on the backward stream
        grid = int(_os.environ.get("APEX_B200_DIST_OVERLAP_CTAS", "64" if mode == 1 else "148"))
        with torch.cuda.stream(self._side_stream):
            self._launch(seg, mode, group, group["step"] if mode == 0 else 1, b, b + 1, grid=grid, done_ctr=self._done_ctr_side, gate=True)
        if mode == 0:
            seg.bucket_stepped[b] = True
        else:
            seg.bucket_synced[b] = True
        self._overlap_launched = True

    def _join_overlap(self):
        if self._overlap_launched:
            torch.cuda.current_stream(self.device).wait_stream(self._side_stream)
            self._overlap_launched = False

    def _join_push(self):
        """Wait for an in-flight parameter push (overlap_param_sync) on the current stream."""
        if self._push_in_flight:
            torch.cuda.current_stream(self.device).wait_stream(self._side_stream)
            self._push_in_flight = False
            for seg in self._segments:
                if seg.region is not None:
                    seg.region.in_flight = False

    def init_param_buffer(self) -> None:
        self.init_params()

    def init_params_bucket(self, params, dtype: Optional[torch.dtype] = None, grad_sync_dtype: Optional[torch.dtype] = None,
                           param_sync_dtype: Optional[torch.dtype] = None, **kwargs) -> None:
        """Reference :1275-1344: initialise ``params`` together (their own bucket), optionally with their own dtypes. The layout here is one
        contiguous space per (group, dtypes) that is cut into buckets afterwards, so there is nothing to place; what is honoured is the
        per-parameter dtype override -- recorded now, applied when the layout is built (first ``step`` / ``zero_grad`` / ``init_params()``).
        Several calls (one per layer, as NeMo / Megatron do) may precede that point: this call never triggers the layout itself."""
        params = [params] if isinstance(params, torch.Tensor) else list(params)
        if self._inited:
            if dtype or grad_sync_dtype or param_sync_dtype:
                raise RuntimeError("init_params_bucket with dtype overrides must be called before the optimizer state is laid out")
            return
        if dtype or grad_sync_dtype or param_sync_dtype:
            self.init_params(params, dtype=dtype, grad_sync_dtype=grad_sync_dtype, param_sync_dtype=param_sync_dtype)

    def set_initial_values(self, param, values):
        """Optional: higher-precision initial values for a low-precision parameter's fp32 master copy."""
        self._init_values[id(param)] = values

    def parameters(self):
        for g in self.param_groups:
            yield from g["params"]

    @dataclass
    class ParameterFragment:
        """One (parameter x bucket) intersection, with the field names of the reference (distributed_fused_adam.py:388-413). Ranges are
        half-open element ranges; ``bucket_id`` counts the buckets of all segments in creation order."""
        param_group_id: int
        param_id: int
        bucket_id: int
        param_range: Tuple[int, int]
        bucket_range: Tuple[int, int]
        in_local_shard: bool
        shard_range: Optional[Tuple[int, int]]
        shard_bucket_range: Optional[Tuple[int, int]]
        shard_param_range: Optional[Tuple[int, int]]

    def parameter(self, *args) -> torch.nn.Parameter:
        """``parameter(param_group_id, param_id)`` or ``parameter(fragment)`` (reference :1199-1226)."""
        if len(args) == 2 and all(isinstance(a, int) for a in args):
            gi, pi = args
        elif len(args) == 1 and isinstance(args[0], self.ParameterFragment):
            gi, pi = args[0].param_group_id, args[0].param_id
        else:
            raise TypeError("Expected input types are [int, int] or [DistributedFusedAdam.ParameterFragment], "
                            f"but found {[type(a).__name__ for a in args]}")
        return self.param_groups[gi]["params"][pi]

    def param_fragments(self, param: torch.nn.Parameter) -> list:
        """The fragments of an initialised parameter — what the reference keeps in ``state[param]["fragments"]`` — computed from the
        segment geometry: a parameter occupies ``[offset, offset + numel)`` of its segment's flat space, bucket ``b`` covers
        ``[b * B, (b + 1) * B)`` of it and rank ``r`` owns ``[r * B / D, (r + 1) * B / D)`` of every bucket."""
        self.init_params()
        si, pi = next((si, pi) for si, sg in enumerate(self._segments) for pi, q in enumerate(sg.params) if q is param)
        seg = self._segments[si]
        base = sum(s.n_buckets for s in self._segments[:si])
        group_params = self.param_groups[seg.group_idx]["params"]
        param_id = next(i for i, q in enumerate(group_params) if q is param)
        B, Sb, r = seg.bucket_elems, seg.shard_elems, seg.rank
        lo, hi = seg.offsets[pi], seg.offsets[pi] + param.numel()
        out = []
        for b in range(lo // B, (max(hi, lo + 1) - 1) // B + 1):
            a, z = max(lo, b * B), min(hi, (b + 1) * B)
            s_lo, s_hi = max(a, b * B + r * Sb), min(z, b * B + (r + 1) * Sb)
            local = s_hi > s_lo
            out.append(self.ParameterFragment(
                param_group_id=seg.group_idx, param_id=param_id, bucket_id=base + b, param_range=(a - lo, z - lo),
                bucket_range=(a - b * B, z - b * B), in_local_shard=local,
                shard_range=(s_lo - b * B - r * Sb, s_hi - b * B - r * Sb) if local else None,
                shard_bucket_range=(s_lo - b * B, s_hi - b * B) if local else None,
                shard_param_range=(s_lo - lo, s_hi - lo) if local else None))
        return out

    # ---------------------------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = False) -> None:
        """``set_to_none=False`` (reference default, :1551-1598): ``p.grad`` becomes a zeroed view of the contiguous gradient buffer and
        backward accumulates into it in place. ``set_to_none=True``: ``p.grad = None``; the buffer is NOT zeroed — every gradient that
        autograd produces is moved into its buffer slot by the post-accumulate hook with a single copy (accumulated from the second
        micro-batch on), which halves the gradient bookkeeping traffic; slots of parameters that got no gradient are zeroed at step()."""
        self.init_params()
        self._join_overlap()
        self._steal = bool(set_to_none) and bool(self._hook_handles)
        for seg in self._segments:
            seg.synced = False
            if self._steal:
                for p in seg.params:
                    p.grad = None
            else:
                seg.grad_buf.zero_()
                seg.attach_grads()
        self._grad_scale.fill_(1.0)
        self._grad_norm = None

    def grad_buffer_view(self, param: torch.nn.Parameter) -> torch.Tensor:
        self.init_params()
        return self._grad_view[id(param)]

    @contextlib.contextmanager
    def no_sync(self, greedy_grad_copy: bool = False):
        """Gradient accumulation: backward passes inside this context only accumulate into the gradient buffer; no bucket is
        reduce-scattered until a backward pass (or step) outside of it (reference :1474-1506)."""
        old = self._sync_enabled
        self._sync_enabled = False
        try:
            yield
        finally:
            self._sync_enabled = old

    def _collect_grads(self):
        """Fold gradients that are not already views of the gradient buffer (user-assigned / dtype-mismatched) into it."""
        for seg in self._segments:
            for pi, p in enumerate(seg.params):
                g = p.grad
                if g is None:
                    if self._steal and not seg.written[pi]:
                        self._grad_view[id(p)].zero_()   # no gradient this step: the slot must not keep the previous step's values
                        seg.written[pi] = True
                    continue
                gv = self._grad_view[id(p)]
                if g is gv or (g.data_ptr() == gv.data_ptr() and g.dtype == gv.dtype):
                    continue
                if any(seg.bucket_synced):
                    raise RuntimeError("a gradient outside the gradient buffer appeared after part of this step's gradients were already "
                                       "reduce-scattered (overlap_grad_sync); assign gradients before backward or disable the overlap")
                if self._steal and not seg.written[pi]:
                    gv.copy_(g.detach())
                    seg.written[pi] = True
                else:
                    gv.add_(g.detach().to(gv.dtype))
                p.grad = None if self._steal else (gv if p.dtype == seg.grad_dtype else None)

    # ---- gradient synchronisation -----------------------------------------------------------------------------------
    def _pre_scale(self):
        return 1.0 / (self.distributed_size * self.redundant_size) if self.average_grad_sync else 1.0

    def _launch(self, seg: _Segment, mode: int, group, step: int, b0: int = 0, b1: Optional[int] = None, grid: Optional[int] = None,
                done_ctr=None, lane: int = 0, force_nvls: Optional[bool] = None, publish: bool = False, gate: bool = False):
        """One csrc/dist_adam.cu launch over buckets [b0, b1) of a segment (default: all of them). ``lane`` selects an independent set
        of signal channels / epoch counter / scratch so that two launches can be in flight at once (hybrid NVLS + P2P step)."""
        b1 = seg.n_buckets if b1 is None else b1
        if mode in (0, 1):
            seg.norm_rows.append(b0)
        D = seg.D
        fused_comm = seg.fused and D > 1
        beta1, beta2 = group["betas"]
        if seg.reduced is None and mode in (1, 2, 3):
            seg.reduced = torch.zeros(seg.local_elems, dtype=torch.float32, device=self.device)
        if fused_comm:
            pad = self._pad
            epoch, epoch_ctr = 0, pad.dev_epochs[2 * lane:2 * lane + 1].data_ptr()   # the epoch lives on the device: graph replays keep counting
            g_arr, p_arr, pads = seg.symm_g.peer_ptr_array(), seg.symm_p.peer_ptr_array(), pad.ptrs
            # NVSwitch multicast moves 16 + 16/D GB per direction per step, plain P2P (D-1)/D * 32 GB: NVLS wins from D = 4 up
            import os as _os

            pol = _os.environ.get("APEX_B200_DIST_NVLS", "auto")
            nvls = int(seg.symm_g.has_multicast and seg.symm_p.has_multicast and (pol == "1" or (pol == "auto" and D >= 4)))
            if force_nvls is not None:
                nvls = int(bool(force_nvls) and seg.symm_g.has_multicast and seg.symm_p.has_multicast)
            mcg, mcp = seg.symm_g.mc_ptr, seg.symm_p.mc_ptr
            self.last_nvls = bool(nvls)
            rank, world = seg.rank, D
        else:
            epoch, epoch_ctr, nvls, mcg, mcp, rank, world = 0, None, 0, 0, 0, 0, 1
            g_arr = (ctypes.c_uint64 * 8)(seg.grad_buf.data_ptr(), 0, 0, 0, 0, 0, 0, 0)
            p_arr = (ctypes.c_uint64 * 8)(seg.param_buf.data_ptr(), 0, 0, 0, 0, 0, 0, 0)
            pads = (ctypes.c_uint64 * 8)(0, 0, 0, 0, 0, 0, 0, 0)
        src_tab, src_n = (seg.source_table() if (D == 1 and mode != 2) else (None, 0))
        if grid is None:
            grid = 148 * (6 if D == 1 else 2)   # measured best per instantiation (csrc/dist_adam.cu: 4 resident CTAs / SM at world size 1, 2 otherwise)
        cap = self.capturable
        done_ctr = self._done_ctr if done_ctr is None else done_ctr
        gate = bool(gate and fused_comm and mode != 2)
        if gate:   # the cross-rank start barrier as a 1-warp kernel in front (launches that overlap with backward)
            _lib.fn("ab_symm_gate")(ctypes.addressof(pads), rank, world, epoch_ctr, 2 * lane, _lib.stream_ptr(self.device))
        self.kernel_launches += 1
        _lib.fn("ab_dist_adam_step")(
            mode, nvls, ctypes.addressof(g_arr), ctypes.addressof(p_arr), ctypes.addressof(pads), mcg, mcp,
            _lib.ptr(seg.master), _lib.ptr(seg.remainders), seg.exp_avg.data_ptr(), seg.exp_avg_sq.data_ptr(), _lib.ptr(seg.reduced), seg.bucket_elems,
            seg.shard_elems, b0, b1, seg.rank, rank, world, epoch, epoch_ctr, 2 * lane, 2 * lane + 1, (seg.group_idx % 32) + 32 * lane,
            done_ctr.data_ptr(), seg.norm_partials[512 * lane:].data_ptr(), seg.norm_out[b0].data_ptr(), self._grad_scale.data_ptr(), self._pre_scale(),
            0.0 if cap else float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), 0 if cap else int(step),
            1 if self.adam_w_mode else 0, 1 if group["bias_correction"] else 0, float(group["weight_decay"]),
            self._dummy_overflow_buf.data_ptr(), group["lr"].data_ptr() if cap else None, group["step"].data_ptr() if cap else None,
            ctypes.addressof(seg.ready_ptrs) if (publish and seg.region is not None) else None,
            seg.bucket_ctr.data_ptr() if (publish and seg.region is not None) else None,
            (seg.region.epoch & 0x7FFFFFFF) if (publish and seg.region is not None) else 0, int(gate), src_tab, src_n,
            _lib.dt(seg.grad_dtype), _lib.dt(seg.param_dtype), grid, _lib.stream_ptr(self.device))

    def _hybrid_split(self, seg: _Segment) -> int:
        """Buckets [0, k) go through the NVSwitch (multimem), [k, n) over plain P2P, as two co-resident kernels: NVLS reduction /
        multicast tops out below the link rate (the switch engines, not the wires, are the limit), so the P2P kernel uses what is left.
        ``APEX_B200_DIST_HYBRID`` = fraction of the buckets on the NVLS side (1 = all NVLS, 0 = all P2P; default 1 below 4 ranks)."""
        import os as _os

        D = seg.D
        if not (seg.fused and D >= 4 and seg.symm_g.has_multicast and seg.symm_p.has_multicast and seg.n_buckets >= 8):
            return -1
        if _os.environ.get("APEX_B200_DIST_NVLS", "auto") not in ("auto",):
            return -1
        f = float(_os.environ.get("APEX_B200_DIST_HYBRID", "1.0"))
        if f >= 1.0 or f <= 0.0:
            return -1
        return max(1, min(seg.n_buckets - 1, int(round(seg.n_buckets * f))))

    def _fused_pass(self, seg: _Segment, group, step):
        """MODE_FUSED over the whole segment: one launch, or the hybrid pair (NVLS lane 0 on the current stream, P2P lane 1 on a forked one)."""
        k = self._hybrid_split(seg)
        if k < 0:
            self._launch(seg, 0, group, step)
            return
        if getattr(self, "_lane_stream", None) is None:
            self._lane_stream = torch.cuda.Stream(device=self.device)
            self._done_ctr_lane = torch.zeros(4, dtype=torch.int32, device=self.device)
        cur = torch.cuda.current_stream(self.device)
        self._lane_stream.wait_stream(cur)
        # one CTA per SM each, launched in the same order on every rank: both grids are resident together, neither can starve the other
        self._launch(seg, 0, group, step, 0, k, grid=148, lane=0, force_nvls=True)
        with torch.cuda.stream(self._lane_stream):
            self._launch(seg, 0, group, step, k, seg.n_buckets, grid=148, done_ctr=self._done_ctr_lane, lane=1, force_nvls=False)
        cur.wait_stream(self._lane_stream)

    def _push_async(self, seg: _Segment, mode: int, group, step):
        """overlap_param_sync: the kernel that updates and pushes the parameters runs on the side stream with a small grid (the
        all-gather is link-bound: a few dozen CTAs saturate it and the rest of the GPU stays free for the next forward); it releases a
        ready flag per (bucket, rank). step() returns at once — consumers synchronise per weight tile / per bucket (parallel/param_sync.py)."""
        import os as _os

        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        self._side_stream.wait_stream(cur)
        seg.region.epoch += 1
        seg.region.in_flight = True
        grid = int(_os.environ.get("APEX_B200_DIST_PARAM_SYNC_CTAS", "48"))
        with torch.cuda.stream(self._side_stream):
            self._launch(seg, mode, group, step, grid=grid, done_ctr=self._done_ctr_side, publish=True)
        self._push_in_flight = True

    def _reduce_scatter_generic(self, seg: _Segment):
        """NCCL / gloo reduce-scatter of every bucket into the fp32 reduced shard (the reference's data path)."""
        D = seg.D
        if seg.reduced is None:
            seg.reduced = torch.zeros(seg.local_elems, dtype=torch.float32, device=self.device)
        red = seg.reduced.view(seg.n_buckets, seg.shard_elems)
        pre = self._pre_scale()
        if D == 1:
            red.copy_(seg.grad_buf.view(seg.n_buckets, seg.shard_elems).float())
            if pre != 1.0:
                red.mul_(pre)
        else:
            pg = self.distributed_process_group
            backend = dist.get_backend(pg)
            for b in range(seg.n_buckets):
                bucket = seg.grad_buf[b * seg.bucket_elems:(b + 1) * seg.bucket_elems]
                if backend == "nccl":
                    out = torch.empty(seg.shard_elems, dtype=seg.grad_dtype, device=self.device)
                    dist.reduce_scatter_tensor(out, bucket, op=dist.ReduceOp.SUM, group=pg)
                    red[b].copy_(out.float())
                else:
                    tmp = bucket.float() if bucket.dtype in (torch.float16, torch.bfloat16) else bucket.clone()
                    dist.all_reduce(tmp, group=pg)
                    red[b].copy_(tmp.view(D, seg.shard_elems)[seg.rank].float())
            if pre != 1.0:
                red.mul_(pre)
        if self.redundant_size > 1:
            dist.all_reduce(seg.reduced, group=self.redundant_process_group)
        sq = (seg.reduced.double() ** 2).sum().float()
        seg.norm_out[0, 0] = sq
        tot = sq.clone()
        if D > 1:
            dist.all_reduce(tot, group=self.distributed_process_group)
        seg.norm_out[0, 1] = tot
        seg.norm_rows = [0]

    def grad_sync(self) -> None:
        """Make sure every segment's reduced shard holds the reduce-scattered gradients of this step."""
        self.init_params()
        self._collect_grads()
        self._join_overlap()
        for seg in self._segments:
            if seg.synced:
                continue
            group = self.param_groups[seg.group_idx]
            if seg.fused:
                self._sync_remaining(seg, group)
            else:
                self._reduce_scatter_generic(seg)
                seg.bucket_synced = [True] * seg.n_buckets

    def _sync_remaining(self, seg: _Segment, group):
        """MODE_RS over every maximal run of buckets that was not reduce-scattered during backward."""
        b = 0
        while b < seg.n_buckets:
            if seg.bucket_synced[b]:
                b += 1
                continue
            e = b
            while e < seg.n_buckets and not seg.bucket_synced[e]:
                e += 1
            self._launch(seg, 1, group, 1, b, e)
            for k in range(b, e):
                seg.bucket_synced[k] = True
            b = e

    def param_sync(self) -> None:
        """Parameters are pushed by the step itself (fused path) or all-gathered inside step() (generic path). With
        ``overlap_param_sync=True`` the push may still be in flight when step() returns: this joins it (flag-aware GEMMs and the
        hooks of ``parallel.param_sync.attach_param_sync_hooks`` do not need it — they wait per tile / per bucket)."""
        self._join_overlap()
        self._join_push()

    def grad_norm(self, parameters=None, norm_type: float = 2.0, force: bool = False) -> torch.Tensor:
        """L2 norm of the (averaged, still loss-scaled) gradients over all ranks; cached until the next zero_grad/step."""
        if norm_type != 2.0:
            raise NotImplementedError("only the L2 norm is supported")
        if self._grad_norm is None or force:
            self.grad_sync()
            tot = torch.zeros([], dtype=torch.float32, device=self.device)
            for seg in self._segments:
                tot = tot + seg.grad_sq()
            self._grad_norm = tot.sqrt()
        return self._grad_norm.detach() * self._grad_scale_for_norm()

    def _grad_scale_for_norm(self):
        return self._grad_scale.detach()

    def clip_grad_norm(self, max_norm: float, parameters=None, norm_type: float = 2.0) -> torch.Tensor:
        """Deferred clipping: folds min(1, max_norm/(norm+1e-6)) into the device ``_grad_scale`` consumed by the step kernel."""
        assert max_norm > 0
        total_norm = self.grad_norm(parameters=parameters, norm_type=norm_type)
        clip_coef = torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)
        self._grad_scale *= clip_coef
        return total_norm

    @torch.no_grad()
    def unscale_grads(self, *args, inv_scale: Optional[torch.Tensor] = None, grad_scaler=None):
        if inv_scale is None and len(args) >= 1:
            inv_scale = args[0]
        found_inf = torch.logical_not(torch.isfinite(self.grad_norm()))
        found_inf_per_device = {found_inf.device: found_inf.float()}
        if grad_scaler is not None and grad_scaler._enabled:
            st = grad_scaler._per_optimizer_states[id(self)]
            OptState = torch.amp.grad_scaler.OptState
            if st["stage"] is OptState.UNSCALED:
                raise RuntimeError("unscale_grads has already been called since the last GradScaler update")
            if st["stage"] is OptState.STEPPED:
                raise RuntimeError("unscale_grads is being called after optimizer step")
            if inv_scale is not None:
                raise ValueError("unscale_grads is being called with both scale_inv and grad_scaler")
            inv_scale = grad_scaler._scale.double().reciprocal().to(dtype=torch.float32, device=self.device)
            st["found_inf_per_device"] = found_inf_per_device
            st["stage"] = OptState.UNSCALED
        if inv_scale is None:
            raise ValueError("unscale_grads is being called with neither scale_inv and grad_scaler")
        self._grad_scale *= inv_scale.view([])
        return found_inf_per_device

    # ---- step -----------------------------------------------------------------------------------------------------------
    def _generic_local_step(self, seg: _Segment, group, step: int):
        """Adam on the local shard with the multi-tensor kernels (any state/grad/param dtype), then all-gather by NCCL/gloo."""
        D = seg.D
        beta1, beta2 = group["betas"]
        sv = seg.shard_view(seg.param_buf)
        inplace = sv.is_contiguous()  # single bucket: the shard is already one contiguous run of the parameter buffer
        out_shard = sv.reshape(-1) if inplace else sv.contiguous().view(-1)
        mode = 1 if self.adam_w_mode else 0
        bc = 1 if group["bias_correction"] else 0
        scaled = seg.scales is not None
        int_sync = not seg.param_dtype.is_floating_point   # integer transport dtype: Adam runs on the master, its top bytes are what is gathered
        if int_sync:
            if seg.master is None:
                raise RuntimeError("an integer param_sync_dtype needs store_params=True (the master holds the real values)")
            int_shard, out_shard = out_shard, seg.master   # the step writes the master in place; packing follows below
        if scaled:
            # 16-bit state with per-fragment scales: widen to fp32 temporaries, step, re-quantise with fresh scales
            st32 = {k: seg.load_scaled(k) for k in ("param", "exp_avg", "exp_avg_sq")}
            real = (seg.master, seg.exp_avg, seg.exp_avg_sq)
            seg.master, seg.exp_avg, seg.exp_avg_sq = st32["param"], st32["exp_avg"], st32["exp_avg_sq"]
        if self.device.type == "cuda":
            g = seg.reduced
            if seg.remainders is not None:
                lists = [[out_shard.view(torch.int16)], [seg.remainders], [seg.exp_avg], [seg.exp_avg_sq], [g], [out_shard.view(torch.int16)]]
                tb = amp_C.TensorTable(lists)
                _lib.fn("ab_mt_dist_adam_remainders")(*tb.head(), 0, self._grad_scale.data_ptr(), float(group["lr"]), float(beta1), float(beta2),
                                                      float(group["eps"]), int(step), mode, bc, float(group["weight_decay"]),
                                                      _lib.stream_ptr(self.device))
            else:
                p_in = seg.master if seg.master is not None else out_shard
                if p_in.dtype != seg.exp_avg.dtype:
                    raise RuntimeError("without store_params the parameter dtype must equal the state dtype")
                tb = amp_C.TensorTable([[p_in], [seg.exp_avg], [seg.exp_avg_sq], [g], [out_shard]])
                d = tb.dtypes
                cap = self.capturable   # device lr / step / overflow flag: no host sync, an overflow step is skipped inside the kernel
                _lib.fn("ab_mt_dist_adam")(*tb.head(), d[0], d[3], d[4], self._grad_scale.data_ptr(), 0.0 if cap else float(group["lr"]),
                                           float(beta1), float(beta2), float(group["eps"]), 0 if cap else int(step), mode, bc,
                                           float(group["weight_decay"]), 1 if cap else 0, group["lr"].data_ptr() if cap else None,
                                           group["step"].data_ptr() if cap else None,
                                           self._dummy_overflow_buf.data_ptr() if cap else None, _lib.stream_ptr(self.device))
        elif seg.remainders is not None:
            # bf16 parameter + int16 remainder ARE the fp32 master: (hi << 16) + lo with a signed lo (hi was rounded to nearest)
            hi, lo = out_shard.view(torch.int16).to(torch.int32), seg.remainders.to(torch.int32)
            master = ((hi << 16) + lo).view(torch.float32).clone()
            ref.dist_adam(master, seg.exp_avg, seg.exp_avg_sq, seg.reduced, None, self._grad_scale, group["lr"], beta1, beta2, group["eps"], step,
                          mode, bc, group["weight_decay"])
            bits = master.view(torch.int32)
            new_lo = ((bits & 0xFFFF) ^ 0x8000) - 0x8000
            seg.remainders.copy_(new_lo.to(torch.int16))
            out_shard.view(torch.int16).copy_(((bits - new_lo) >> 16).to(torch.int16))
        else:
            p_in = seg.master if seg.master is not None else out_shard
            ref.dist_adam(p_in, seg.exp_avg, seg.exp_avg_sq, seg.reduced, out_shard, self._grad_scale, group["lr"], beta1, beta2,
                          group["eps"], step, mode, bc, group["weight_decay"])
        if int_sync:
            _pack_msb(seg.master, int_shard)
            out_shard = int_shard
        if scaled:
            seg.master, seg.exp_avg, seg.exp_avg_sq = real
            for k in ("param", "exp_avg", "exp_avg_sq"):
                seg.store_scaled(k, st32[k])
        # scatter the contiguous shard back into the bucket-interleaved parameter buffer and all-gather
        if not inplace:
            sv.copy_(out_shard.view(seg.n_buckets, seg.shard_elems))
        if D > 1:
            pg = self.distributed_process_group
            for b in range(seg.n_buckets):
                bucket = seg.param_buf[b * seg.bucket_elems:(b + 1) * seg.bucket_elems]
                mine = bucket[seg.rank * seg.shard_elems:(seg.rank + 1) * seg.shard_elems]
                if dist.get_backend(pg) == "nccl":
                    dist.all_gather_into_tensor(bucket, mine.clone(), group=pg)
                else:
                    parts = [torch.empty_like(mine) for _ in range(D)]
                    dist.all_gather(parts, mine.clone(), group=pg)
                    bucket.copy_(torch.cat(parts))

    def shard_view_contig(self, seg: _Segment) -> torch.Tensor:
        """Contiguous copy of this rank's low-precision parameter shard (generic path scratch)."""
        return seg.shard_view(seg.param_buf).contiguous().view(-1)

    @torch.no_grad()
    def step(self, closure=None, *, grad_scaler=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.init_params()
        self._collect_grads()
        self._join_overlap()
        self._join_push()     # the previous step's parameter push (overlap_param_sync), if a consumer has not joined it already

        if any(any(s.bucket_stepped) for s in self._segments):
            if grad_scaler is not None or self._grad_norm is not None:
                raise RuntimeError("overlap_step_with_backward already applied part of this step during backward: it cannot be combined "
                                   "with clip_grad_norm / grad_norm / GradScaler (construct the optimizer without it)")
            for seg in self._segments:   # whatever backward did not reach (parameters without a hook firing) steps now
                group = self.param_groups[seg.group_idx]
                b = 0
                while seg.fused and b < seg.n_buckets:
                    if seg.bucket_stepped[b]:
                        b += 1
                        continue
                    e = b
                    while e < seg.n_buckets and not seg.bucket_stepped[e]:
                        e += 1
                    self._launch(seg, 0, group, group["step"], b, e)
                    b = e
            self._finish_step(skipped=False)
            return loss

        need_two_phase = grad_scaler is not None or self._grad_norm is not None or any(any(s.bucket_synced) for s in self._segments)
        if grad_scaler is not None:
            st = grad_scaler._per_optimizer_states[id(self)]
            if st["stage"] is not torch.amp.grad_scaler.OptState.UNSCALED:
                self.unscale_grads(grad_scaler=grad_scaler)
            found = sum(v.to(self.device) for v in st["found_inf_per_device"].values())
            self._dummy_overflow_buf.copy_((found > 0).to(torch.int32).reshape(1))
            # kernels that take the device flag skip the update themselves; the remainder kernel and the CPU routines do not,
            # so those configurations read the flag on the host even when capturable
            host_skip = (not self.capturable) or any((not seg.fused) and (seg.remainders is not None or self.device.type != "cuda")
                                                     for seg in self._segments)
            if host_skip and int(self._dummy_overflow_buf.item()) != 0:
                self._finish_step(skipped=True)
                return loss
        else:
            self._dummy_overflow_buf.zero_()

        self._bump_step()
        for seg in self._segments:
            group = self.param_groups[seg.group_idx]
            step = group["step"]
            if seg.fused:
                if seg.region is not None:
                    # overlap_param_sync: reduce-scatter now (whatever backward did not already overlap), then Adam + push on the side
                    # stream; the gradient buffer is free again as soon as the reduce-scatter is done, so the next zero_grad / backward
                    # does not have to wait for the parameter push
                    if not seg.synced:
                        self._sync_remaining(seg, group)
                    self._push_async(seg, 2, group, step)
                    continue
                if need_two_phase and not seg.synced:
                    self._sync_remaining(seg, group)
                if need_two_phase:
                    self._launch(seg, 2, group, step)
                else:
                    self._fused_pass(seg, group, step)
            else:
                if not seg.synced:
                    self._reduce_scatter_generic(seg)
                self._generic_local_step(seg, group, float(step) if (self.capturable and self.device.type != "cuda") else step)
        self._finish_step(skipped=False)
        return loss

    def last_grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the gradients consumed by the most recent step — a by-product of the fused kernel (no extra pass,
        no extra collective). Device tensor; reading it on the host is the only synchronisation."""
        tot = torch.zeros([], dtype=torch.float32, device=self.device)
        for seg in self._segments:
            tot = tot + seg.norm_out[self._last_norm_rows.get(id(seg), [0]), 1].sum()
        return tot.sqrt()

    def _finish_step(self, skipped: bool):
        self._step_bumped = False
        self._last_norm_rows = {id(seg): list(seg.norm_rows) or [0] for seg in self._segments}
        if any(seg.cast_params and seg.region is not None and seg.region.in_flight for seg in self._segments):
            self._join_push()   # the cast copies below read the gathered buffer
        for seg in self._segments:
            seg.synced = False
            for p in seg.cast_params:  # parameters whose dtype differs from the sync dtype get a cast copy
                if not seg.param_dtype.is_floating_point:
                    _pack_msb(self._param_view[id(p)], p.data)   # the same byte transport in the other direction
                elif p.dtype != seg.param_dtype:
                    p.data.copy_(self._param_view[id(p)].to(p.dtype))
        self._grad_scale.fill_(1.0)
        self._grad_norm = None

    # ---- checkpointing: world-size independent (v2-style) -----------------------------------------------------------------
    _CKPT_CHUNK_BYTES = 256 << 20   # full-size staging per gathered piece: the GPU never holds more than this of a gathered state

    def _gather_full(self, seg: _Segment, shard: torch.Tensor) -> torch.Tensor:
        """[local_elems] shard -> full flat [padded] tensor (identical on every rank). Small layouts / tests only; checkpoints stream
        through :meth:`_stream_full`."""
        pieces = []
        self._stream_full(seg, shard, lambda start, full: pieces.append(full.clone()))
        return torch.cat(pieces)

    def _stream_full(self, seg: _Segment, shard: torch.Tensor, consume) -> None:
        """Walk a sharded state bucket-group by bucket-group: all-gather at most ``_CKPT_CHUNK_BYTES`` of the full tensor, hand
        ``consume(flat_start, full_piece)`` the gathered piece, drop it, continue (reference :3226-3321 streams bucket by bucket too)."""
        D = seg.D
        sh = shard.view(seg.n_buckets, seg.shard_elems)
        per = max(1, self._CKPT_CHUNK_BYTES // max(1, seg.bucket_elems * shard.element_size()))
        for b0 in range(0, seg.n_buckets, per):
            b1 = min(seg.n_buckets, b0 + per)
            piece = sh[b0:b1]
            if D == 1:
                full = piece.reshape(-1)
            else:
                raw = piece.contiguous().view(torch.uint8)   # a pure data move: bytes work for every dtype on every backend (NCCL and gloo lack int16)
                parts = [torch.empty_like(raw) for _ in range(D)]
                dist.all_gather(parts, raw, group=self.distributed_process_group)
                full = torch.stack([q.view(shard.dtype).view(b1 - b0, seg.shard_elems) for q in parts], dim=1).reshape(-1)
            consume(b0 * seg.bucket_elems, full)

    def _param_spans(self, seg: _Segment, start: int, n: int):
        """(param index, offset inside the parameter, offset inside the piece, length) of every parameter intersecting [start, start+n)."""
        import bisect

        out = []
        i = max(0, bisect.bisect_right(seg.offsets, start) - 1)
        end = start + n
        while i < len(seg.params) and seg.offsets[i] < end:
            lo, hi = seg.offsets[i], seg.offsets[i] + seg.params[i].numel()
            a, z = max(lo, start), min(hi, end)
            if z > a:
                out.append((i, a - lo, a - start, z - a))
            i += 1
        return out

    def _global_step(self):
        st = self.param_groups[0].get("step", 0) if self.param_groups else 0
        return int(st.item()) if torch.is_tensor(st) else int(st)

    def _state_dict_v1(self, gather_on_root: bool = True):
        """Deprecated v1 format of the reference (:2907-3057): every rank serialises ITS OWN shard of the optimizer state
        (``torch.save`` into bytes); the root rank returns ``{"gathered_states": [uint8 tensor per rank]}``, the other ranks ``None``.
        Tied to the world size and bucket layout it was written with (load_state_dict checks both); v2 is the portable format."""
        import io
        import warnings

        import torch.distributed as dist

        warnings.warn("Making optimizer state dictionary in deprecated v1 format. Future support is not guaranteed.")
        self.init_params()
        self._join_overlap()
        self._join_push()
        if any(seg.scales is not None for seg in self._segments):
            raise NotImplementedError("Deprecated v1 format does not support scaled state")
        local = {"world": self.distributed_size, "rank": self.distributed_rank, "step": self._global_step(),
                 "param_groups": [{k: (v.item() if torch.is_tensor(v) and v.numel() == 1 else v) for k, v in g.items() if k != "params"}
                                  for g in self.param_groups],
                 "segments": []}
        for seg in self._segments:
            ent = {"local_elems": int(seg.local_elems), "n_params": len(seg.params), "exp_avg": seg.exp_avg.detach().cpu(),
                   "exp_avg_sq": seg.exp_avg_sq.detach().cpu()}
            if seg.master is not None:
                ent["master"] = seg.master.detach().cpu()
            if seg.remainders is not None:
                ent["remainders"] = seg.remainders.detach().cpu()
            local["segments"].append(ent)
        if not gather_on_root:
            return local
        buf = io.BytesIO()
        torch.save(local, buf)
        mine = torch.frombuffer(bytearray(buf.getvalue()), dtype=torch.uint8)
        if self.distributed_size == 1:
            return {"gathered_states": [mine], "format": 1}
        gathered = [None] * self.distributed_size if self.distributed_rank == 0 else None
        dist.gather_object(mine, gathered, dst=dist.get_global_rank(self.distributed_process_group, 0) if self.distributed_process_group is not None else 0,
                           group=self.distributed_process_group)
        return {"gathered_states": gathered, "format": 1} if self.distributed_rank == 0 else None

    def _load_state_dict_v1(self, state_dict) -> None:
        import io

        import torch.distributed as dist

        self.init_params()
        self._join_overlap()
        self._join_push()
        if self.distributed_size > 1:   # the root holds everybody's bytes: hand each rank its own
            mine = [None]
            src = dist.get_global_rank(self.distributed_process_group, 0) if self.distributed_process_group is not None else 0
            dist.scatter_object_list(mine, state_dict["gathered_states"] if self.distributed_rank == 0 else None, src=src, group=self.distributed_process_group)
            blob = mine[0]
        else:
            blob = state_dict["gathered_states"][0]
        local = torch.load(io.BytesIO(bytes(blob.numpy().tobytes())), weights_only=False)
        if local["world"] != self.distributed_size or len(local["segments"]) != len(self._segments):
            raise ValueError(f"v1 optimizer state was written by {local['world']} ranks / {len(local['segments'])} segments; this optimizer has "
                             f"{self.distributed_size} / {len(self._segments)} (the v1 format cannot be resharded: save in the default v2 format)")
        for seg, ent in zip(self._segments, local["segments"]):
            if ent["local_elems"] != int(seg.local_elems) or ent["n_params"] != len(seg.params):
                raise ValueError("v1 optimizer state does not match this optimizer's bucket layout")
            seg.exp_avg.copy_(ent["exp_avg"])
            seg.exp_avg_sq.copy_(ent["exp_avg_sq"])
            if seg.master is not None and "master" in ent:
                seg.master.copy_(ent["master"])
            if seg.remainders is not None and "remainders" in ent:
                seg.remainders.copy_(ent["remainders"])
        for g, sg in zip(self.param_groups, local["param_groups"]):
            for k, v in sg.items():
                if self.capturable and k in ("lr", "step"):
                    g[k].copy_(torch.as_tensor(v).reshape(1).to(g[k].dtype))
                else:
                    g[k] = v
        for g in self.param_groups:
            if self.capturable:
                g["step"].fill_(int(local["step"]))
            else:
                g["step"] = int(local["step"])
        # as in the reference, the model-dtype parameters themselves come from the MODEL's checkpoint; v1 restores the optimizer's shards
        # (moments, fp32 master / remainder shards, step, hyper-parameters)

    def state_dict(self, *args, state_dict_format=None, gather_on_root: Optional[bool] = None, **kwargs):
        """``state_dict_format=1``: the deprecated per-rank format (see :meth:`_state_dict_v1`). Default (2):
        every rank returns the same dict in the reference's v2 layout (:3059-3327): ``state["step"]`` plus, per parameter index,
        full-size CPU tensors ``param`` (fp32 master), ``exp_avg``, ``exp_avg_sq`` — independent of world size and bucket layout, so it
        reloads under a different parallel configuration (and in the reference). The sharded state is streamed to the host in
        bounded pieces (double-buffered pinned staging), never materialised at full size on the GPU."""
        if state_dict_format == 1:
            return self._state_dict_v1(True if gather_on_root is None else gather_on_root)
        if state_dict_format not in (None, 2):
            raise ValueError(f"Unrecognized state dict format ({state_dict_format})")
        self.init_params()
        self._join_overlap()
        self._join_push()
        index = {}
        i = 0
        for g in self.param_groups:
            for p in g["params"]:
                index[id(p)] = i
                i += 1
        step = self._global_step()
        state = {"step": step}
        cuda = self.device.type == "cuda" and torch.cuda.is_available()
        for seg in self._segments:
            out_dtype = torch.float32 if seg.scales is not None else seg.dtype
            ents = []
            for p in seg.params:
                ent = {k: torch.zeros(p.shape, dtype=out_dtype) for k in ("param", "exp_avg", "exp_avg_sq")}
                ent["step"] = step   # extra key (torch.optim style); the reference reads state["step"]
                ents.append(ent)
                state[index[id(p)]] = ent
            stage = [None, None]     # pinned double buffer: the D2H copy of piece i overlaps the gather of piece i + 1
            events = [None, None]
            turn = [0]

            def to_host(key, conv=None):
                def consume(start, full):
                    if conv is not None:
                        full = conv(start, full)
                    k = turn[0] & 1
                    turn[0] += 1
                    if cuda:
                        if events[k] is not None:
                            events[k].synchronize()
                            self._drain(stage[k])
                        if stage[k] is None or stage[k][0].numel() < full.numel() or stage[k][0].dtype != full.dtype:
                            stage[k] = [torch.empty(full.numel(), dtype=full.dtype).pin_memory(), None]
                        host = stage[k][0][:full.numel()]
                        host.copy_(full, non_blocking=True)
                        events[k] = torch.cuda.Event()
                        events[k].record()
                        stage[k][1] = (host, [(ents[pi][key], po, fo, ln) for pi, po, fo, ln in self._param_spans(seg, start, full.numel())])
                    else:
                        for pi, po, fo, ln in self._param_spans(seg, start, full.numel()):
                            ents[pi][key].view(-1)[po:po + ln].copy_(full[fo:fo + ln])
                return consume

            def flush():
                for k in (0, 1):
                    if events[k] is not None:
                        events[k].synchronize()
                        self._drain(stage[k])
                        events[k] = None

            for key, t in (("exp_avg", seg.exp_avg), ("exp_avg_sq", seg.exp_avg_sq)):
                self._stream_full(seg, seg.load_scaled(key) if seg.scales is not None else t, to_host(key))
                flush()
            if seg.scales is not None:
                self._stream_full(seg, seg.load_scaled("param"), to_host("param"))
            elif seg.master is not None:
                self._stream_full(seg, seg.master, to_host("param"))
            elif seg.remainders is not None:
                # the checkpoint holds the exact fp32 master: (bf16 bits << 16) + signed remainder (reference :3472-3477 does the same)
                def rebuild(start, lo):
                    hi = seg.param_buf[start:start + lo.numel()].view(torch.int16).to(torch.int32)
                    return ((hi << 16) + lo.to(torch.int32)).view(torch.float32)
                self._stream_full(seg, seg.remainders, to_host("param", rebuild))
            else:
                for p, ent in zip(seg.params, ents):
                    ent["param"].copy_(self._param_view[id(p)].detach().float())
            flush()
        groups = []
        for g in self.param_groups:
            gg = {k: (v.item() if torch.is_tensor(v) and v.numel() == 1 else v) for k, v in g.items() if k != "params"}
            gg["params"] = [index[id(p)] for p in g["params"]]
            groups.append(gg)
        return {"state": state, "param_groups": groups, "format": 2}

    @staticmethod
    def _drain(slot):
        if slot is None or slot[1] is None:
            return
        host, spans = slot[1]
        for dst, po, fo, ln in spans:
            dst.view(-1)[po:po + ln].copy_(host[fo:fo + ln])
        slot[1] = None

    def load_state_dict(self, state_dict) -> None:
        if state_dict is not None and "gathered_states" in state_dict or (state_dict is None and self.distributed_size > 1):
            return self._load_state_dict_v1(state_dict)   # non-root ranks pass None, as they received it from state_dict()
        self.init_params()
        self._join_overlap()
        self._join_push()
        index = {}
        i = 0
        st = state_dict["state"]
        for g, sg in zip(self.param_groups, state_dict["param_groups"]):
            for k, v in sg.items():
                if k == "params":
                    continue
                if self.capturable and k in ("lr", "step"):
                    # the kernels hold the ADDRESS of these device tensors (and so does any captured graph): update in place
                    g[k].copy_(torch.as_tensor(v).reshape(1).to(g[k].dtype))
                elif k in ("lr", "step") and torch.is_tensor(v):
                    g[k] = v.item()
                else:
                    g[k] = v
            for p in g["params"]:
                index[id(p)] = i
                i += 1
        # the reference keeps ONE step counter in state["step"] (:3081, :3427); per-group / per-parameter values are extras of this format
        step = st.get("step")
        if step is None:
            any_ent = next((e for e in st.values() if isinstance(e, dict) and "step" in e), None)
            step = any_ent["step"] if any_ent is not None else None
        if step is not None:
            step = int(step.item()) if torch.is_tensor(step) else int(step)
            for g in self.param_groups:
                if self.capturable:
                    g["step"].fill_(step)
                else:
                    g["step"] = step
        for seg in self._segments:
            Sb, B, r = seg.shard_elems, seg.bucket_elems, seg.rank
            if seg.scales is not None:
                tmp = {k: torch.zeros(seg.local_elems, dtype=torch.float32, device=self.device) for k in ("exp_avg", "exp_avg_sq", "param")}
            per = max(1, self._CKPT_CHUNK_BYTES // max(1, B * 4))
            for b0 in range(0, seg.n_buckets, per):     # bounded staging: one group of buckets of the full fp32 tensor at a time
                b1 = min(seg.n_buckets, b0 + per)
                start, n = b0 * B, (b1 - b0) * B
                spans = self._param_spans(seg, start, n)
                for key in ("exp_avg", "exp_avg_sq", "param"):
                    full = torch.zeros(n, dtype=torch.float32, device=self.device)
                    for pi, po, fo, ln in spans:
                        src = st[index[id(seg.params[pi])]][key].reshape(-1)[po:po + ln]
                        full[fo:fo + ln].copy_(src.to(self.device, torch.float32, non_blocking=True))
                    mine = full.view(b1 - b0, seg.D, Sb)[:, r, :]
                    if seg.scales is not None:
                        tmp[key].view(seg.n_buckets, Sb)[b0:b1].copy_(mine)
                    elif key != "param":
                        getattr(seg, key).view(seg.n_buckets, Sb)[b0:b1].copy_(mine)
                    else:
                        if seg.master is not None:
                            seg.master.view(seg.n_buckets, Sb)[b0:b1].copy_(mine)
                        if seg.remainders is not None:
                            # split the fp32 master with the SAME convention as the step kernel: signed low half, high half = (bits - lo) >> 16
                            bits = full.view(torch.int32)
                            lo = ((bits & 0xFFFF) ^ 0x8000) - 0x8000
                            seg.remainders.view(seg.n_buckets, Sb)[b0:b1].copy_(lo.view(b1 - b0, seg.D, Sb)[:, r, :].to(torch.int16))
                            seg.param_buf[start:start + n].view(torch.int16).copy_(((bits - lo) >> 16).to(torch.int16))
                        elif not seg.param_dtype.is_floating_point:
                            _pack_msb(full.to(seg.dtype), seg.param_buf[start:start + n])
                        else:
                            seg.param_buf[start:start + n].copy_(full.to(seg.param_dtype))
                    del full
            if seg.scales is not None:
                for k in ("exp_avg", "exp_avg_sq", "param"):
                    seg.store_scaled(k, tmp[k])
            for p in seg.cast_params:
                if not seg.param_dtype.is_floating_point:
                    _pack_msb(self._param_view[id(p)], p.data)
                elif p.dtype != seg.param_dtype:
                    p.data.copy_(self._param_view[id(p)].to(p.dtype))

    def __repr__(self):
        return (f"{type(self).__name__}(distributed_size={self.distributed_size}, redundant_size={self.redundant_size}, "
                f"dtype={self.dtype}, grad_sync_dtype={self._grad_sync_dtype}, param_sync_dtype={self._param_sync_dtype}, "
                f"bucket_cap_mb={self.bucket_cap_mb}, fused_collectives={self.fused_collectives}, segments={len(self._segments)})")


def _smoke(dev):
    """Tiny single-rank step through the fused kernel (world size 1): used by __graft_entry__.smoke()."""
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(1000, 33, device=dev, dtype=torch.bfloat16)), torch.nn.Parameter(torch.randn(77, device=dev, dtype=torch.bfloat16))]
    opt = DistributedFusedAdam(ps, lr=1e-2, bucket_cap_mb=0.1)
    opt.zero_grad()
    before = [p.detach().clone() for p in ps]
    for p in ps:
        p.grad.copy_(torch.randn_like(p))
    opt.step()
    torch.cuda.synchronize()
    assert all((a.float() - b.float()).abs().max() > 0 for a, b in zip(before, ps))
