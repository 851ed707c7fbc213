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
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from ... import _lib
from ...ops import amp_C
from ...ops import reference as ref

_lib.declare("ab_dist_adam_step", "i i p p p l l p p p p l i i i i i i i i i i p p p p f f f f f i i i f p p p i i i p")

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
        off, self.offsets = 0, []
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
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
        self.norm_out = torch.zeros(2, dtype=torch.float32, device=dev)
        self.norm_partials = torch.zeros(1024, dtype=torch.float32, device=dev) if dev.type == "cuda" else None
        self.synced = False      # reduced shard holds this step's reduce-scattered grads
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
                 nccl_ub: bool = False, capturable: bool = False, fused_collectives="auto"):
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
        self.last_nvls = False
        self.kernel_launches = 0  # number of csrc/dist_adam.cu launches so far (bench.py reports it)
        self._last_grad_norm = None
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
        return (self.fused_collectives and not self.with_scaled_states and dtype == torch.float32 and self.store_params and grad_dtype in f and param_dtype in f
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
        self._inited = True
        self._collect_grads()  # gradients that already exist (assigned before the first step) are folded in, not dropped
        for seg in self._segments:
            seg.attach_grads()

    def init_param_buffer(self) -> None:
        self.init_params()

    def init_params_bucket(self, params, **kwargs) -> None:
        """Accepted for API compatibility: the layout here is already one contiguous space per (group, dtypes)."""
        self.init_params()

    def set_initial_values(self, param, values):
        """Optional: higher-precision initial values for a low-precision parameter's fp32 master copy."""
        self._init_values[id(param)] = values

    def parameters(self):
        for g in self.param_groups:
            yield from g["params"]

    # ---------------------------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = False) -> None:
        self.init_params()
        for seg in self._segments:
            seg.grad_buf.zero_()
            seg.synced = False
            seg.attach_grads()
        self._grad_scale.fill_(1.0)
        self._grad_norm = None

    def grad_buffer_view(self, param: torch.nn.Parameter) -> torch.Tensor:
        self.init_params()
        return self._grad_view[id(param)]

    @contextlib.contextmanager
    def no_sync(self, greedy_grad_copy: bool = False):
        old = self._sync_enabled
        self._sync_enabled = False
        try:
            yield
        finally:
            self._sync_enabled = old

    def _collect_grads(self):
        """Fold gradients that are not already views of the gradient buffer (user-assigned / dtype-mismatched) into it."""
        for seg in self._segments:
            for p in seg.params:
                gv = self._grad_view[id(p)]
                g = p.grad
                if g is None or (g.data_ptr() == gv.data_ptr() and g.dtype == gv.dtype):
                    continue
                gv.add_(g.detach().to(gv.dtype))
                p.grad = gv if p.dtype == seg.grad_dtype else None
                seg.synced = False

    # ---- gradient synchronisation -----------------------------------------------------------------------------------
    def _pre_scale(self):
        return 1.0 / (self.distributed_size * self.redundant_size) if self.average_grad_sync else 1.0

    def _launch(self, seg: _Segment, mode: int, group, step: int):
        """One csrc/dist_adam.cu launch over every bucket of a segment."""
        D = seg.D
        fused_comm = seg.fused and D > 1
        beta1, beta2 = group["betas"]
        if seg.reduced is None and mode in (1, 2):
            seg.reduced = torch.zeros(seg.local_elems, dtype=torch.float32, device=self.device)
        if fused_comm:
            pad = self._pad
            epoch = pad.next_epoch()
            g_arr, p_arr, pads = seg.symm_g.peer_ptr_array(), seg.symm_p.peer_ptr_array(), pad.ptrs
            # NVSwitch multicast moves 16 + 16/D GB per direction per step, plain P2P (D-1)/D * 32 GB: NVLS wins from D = 4 up
            import os as _os

            pol = _os.environ.get("APEX_B200_DIST_NVLS", "auto")
            nvls = int(seg.symm_g.has_multicast and seg.symm_p.has_multicast and (pol == "1" or (pol == "auto" and D >= 4)))
            mcg, mcp = seg.symm_g.mc_ptr, seg.symm_p.mc_ptr
            self.last_nvls = bool(nvls)
            rank, world = seg.rank, D
        else:
            epoch, nvls, mcg, mcp, rank, world = 0, 0, 0, 0, 0, 1
            g_arr = (ctypes.c_uint64 * 8)(seg.grad_buf.data_ptr(), 0, 0, 0, 0, 0, 0, 0)
            p_arr = (ctypes.c_uint64 * 8)(seg.param_buf.data_ptr(), 0, 0, 0, 0, 0, 0, 0)
            pads = (ctypes.c_uint64 * 8)(0, 0, 0, 0, 0, 0, 0, 0)
        grid = 148 * (3 if D <= 2 else 2)
        cap = self.capturable
        self.kernel_launches += 1
        _lib.fn("ab_dist_adam_step")(
            mode, nvls, ctypes.addressof(g_arr), ctypes.addressof(p_arr), ctypes.addressof(pads), mcg, mcp,
            _lib.ptr(seg.master), seg.exp_avg.data_ptr(), seg.exp_avg_sq.data_ptr(), _lib.ptr(seg.reduced), seg.bucket_elems,
            seg.shard_elems, 0, seg.n_buckets, seg.rank, rank, world, epoch, 0, 1, seg.group_idx % 64, self._done_ctr.data_ptr(),
            seg.norm_partials.data_ptr(), seg.norm_out.data_ptr(), self._grad_scale.data_ptr(), self._pre_scale(),
            0.0 if cap else float(group["lr"]), float(beta1), float(beta2), float(group["eps"]), 0 if cap else int(step),
            1 if self.adam_w_mode else 0, 1 if group["bias_correction"] else 0, float(group["weight_decay"]),
            self._dummy_overflow_buf.data_ptr(), group["lr"].data_ptr() if cap else None, group["step"].data_ptr() if cap else None,
            _lib.dt(seg.grad_dtype), _lib.dt(seg.param_dtype), grid, _lib.stream_ptr(self.device))

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
        seg.norm_out[0] = sq
        tot = sq.clone()
        if D > 1:
            dist.all_reduce(tot, group=self.distributed_process_group)
        seg.norm_out[1] = tot

    def grad_sync(self) -> None:
        """Make sure every segment's reduced shard holds the reduce-scattered gradients of this step."""
        self.init_params()
        self._collect_grads()
        for seg in self._segments:
            if seg.synced:
                continue
            group = self.param_groups[seg.group_idx]
            if seg.fused:
                self._launch(seg, 1, group, 1)
            else:
                self._reduce_scatter_generic(seg)
            seg.synced = True

    def param_sync(self) -> None:
        """Parameters are pushed by the step itself on the fused path; the generic path all-gathers inside step()."""
        return

    def grad_norm(self, parameters=None, norm_type: float = 2.0, force: bool = False) -> torch.Tensor:
        """L2 norm of the (averaged, still loss-scaled) gradients over all ranks; cached until the next zero_grad/step."""
        if norm_type != 2.0:
            raise NotImplementedError("only the L2 norm is supported")
        if self._grad_norm is None or force:
            self.grad_sync()
            tot = torch.zeros([], dtype=torch.float32, device=self.device)
            for seg in self._segments:
                tot = tot + seg.norm_out[1]
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
                _lib.fn("ab_mt_dist_adam")(*tb.head(), d[0], d[3], d[4], self._grad_scale.data_ptr(), float(group["lr"]), float(beta1),
                                           float(beta2), float(group["eps"]), int(step), mode, bc, float(group["weight_decay"]), 0, None,
                                           None, None, _lib.stream_ptr(self.device))
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

        need_two_phase = grad_scaler is not None or self._grad_norm is not None or any(s.synced for s in self._segments)
        if grad_scaler is not None:
            st = grad_scaler._per_optimizer_states[id(self)]
            if st["stage"] is not torch.amp.grad_scaler.OptState.UNSCALED:
                self.unscale_grads(grad_scaler=grad_scaler)
            found = sum(v.to(self.device) for v in st["found_inf_per_device"].values())
            self._dummy_overflow_buf.copy_((found > 0).to(torch.int32).reshape(1))
            if not self.capturable and int(self._dummy_overflow_buf.item()) != 0:
                self._finish_step(skipped=True)
                return loss
        else:
            self._dummy_overflow_buf.zero_()

        for gi, group in enumerate(self.param_groups):
            if self.capturable:
                group["step"] += (self._dummy_overflow_buf != 1).to(torch.int32)
            else:
                group["step"] = group.get("step", 0) + 1
        for seg in self._segments:
            group = self.param_groups[seg.group_idx]
            step = group["step"]
            if seg.fused:
                if need_two_phase:
                    if not seg.synced:
                        self._launch(seg, 1, group, step)
                    self._launch(seg, 2, group, step)
                else:
                    self._launch(seg, 0, group, step)
            else:
                if not seg.synced:
                    self._reduce_scatter_generic(seg)
                self._generic_local_step(seg, group, step)
        self._finish_step(skipped=False)
        return loss

    def last_grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the gradients consumed by the most recent step — a by-product of the fused kernel (no extra pass,
        no extra collective). Device tensor; reading it on the host is the only synchronisation."""
        tot = torch.zeros([], dtype=torch.float32, device=self.device)
        for seg in self._segments:
            tot = tot + seg.norm_out[1]
        return tot.sqrt()

    def _finish_step(self, skipped: bool):
        for seg in self._segments:
            seg.synced = False
            for p in seg.params:  # parameters whose dtype differs from the sync dtype get a cast copy
                if not seg.param_dtype.is_floating_point:
                    _pack_msb(self._param_view[id(p)], p.data)   # the same byte transport in the other direction
                elif p.dtype != seg.param_dtype:
                    p.data.copy_(self._param_view[id(p)].to(p.dtype))
        self._grad_scale.fill_(1.0)
        self._grad_norm = None

    # ---- checkpointing: world-size independent (v2-style) -----------------------------------------------------------------
    def _gather_full(self, seg: _Segment, shard: torch.Tensor) -> torch.Tensor:
        """[local_elems] shard -> full flat [padded] tensor (identical on every rank)."""
        D = seg.D
        sh = shard.view(seg.n_buckets, seg.shard_elems)
        if D == 1:
            return sh.reshape(-1).clone()
        raw = sh.contiguous().view(torch.uint8)   # a pure data move: bytes work for every dtype on every backend (NCCL and gloo lack int16)
        parts = [torch.empty_like(raw) for _ in range(D)]
        dist.all_gather(parts, raw, group=self.distributed_process_group)
        return torch.stack([p.view(shard.dtype) for p in parts], dim=1).reshape(-1)

    def state_dict(self, *args, **kwargs):
        """Every rank returns the same dict: per-parameter full-size CPU tensors (param master, exp_avg, exp_avg_sq), independent
        of world size and bucket layout, so it can be reloaded under a different parallel configuration (reference v2 format,
        :3059-3327)."""
        self.init_params()
        index = {}
        i = 0
        for g in self.param_groups:
            for p in g["params"]:
                index[id(p)] = i
                i += 1
        state = {}
        for seg in self._segments:
            if seg.scales is not None:
                fulls = {k: self._gather_full(seg, seg.load_scaled(k)) for k in ("exp_avg", "exp_avg_sq", "param")}
            else:
                fulls = {"exp_avg": self._gather_full(seg, seg.exp_avg), "exp_avg_sq": self._gather_full(seg, seg.exp_avg_sq)}
            if seg.master is not None and seg.scales is None:
                fulls["param"] = self._gather_full(seg, seg.master)
            elif seg.remainders is not None:
                fulls["param_remainder"] = self._gather_full(seg, seg.remainders)
            for p, off in zip(seg.params, seg.offsets):
                n = p.numel()
                ent = {k: v[off:off + n].view(p.shape).cpu() for k, v in fulls.items()}
                if "param_remainder" in ent:
                    # the checkpoint holds the exact fp32 master: (bf16 bits << 16) + signed remainder (reference :3472-3477 does the same)
                    hi = seg.param_buf[off:off + n].view(torch.int16).to(torch.int32)
                    lo = fulls["param_remainder"][off:off + n].to(torch.int32)
                    ent["param"] = ((hi << 16) + lo).view(torch.float32).view(p.shape).cpu()
                elif "param" not in ent:
                    ent["param"] = self._param_view[id(p)].detach().float().cpu()
                ent["step"] = self.param_groups[seg.group_idx].get("step", 0)
                state[index[id(p)]] = ent
        groups = []
        for g in self.param_groups:
            gg = {k: v for k, v in g.items() if k != "params"}
            gg["params"] = [index[id(p)] for p in g["params"]]
            groups.append(gg)
        return {"state": state, "param_groups": groups, "format": 2}

    def load_state_dict(self, state_dict) -> None:
        self.init_params()
        index = {}
        i = 0
        for g, sg in zip(self.param_groups, state_dict["param_groups"]):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v
            for p in g["params"]:
                index[id(p)] = i
                i += 1
        for seg in self._segments:
            fulls = {k: torch.zeros(seg.padded, dtype=torch.float32 if seg.scales is not None else t.dtype, device=self.device)
                     for k, t in (("exp_avg", seg.exp_avg), ("exp_avg_sq", seg.exp_avg_sq), ("param", seg.master)) if t is not None}
            pfull = torch.zeros(seg.padded, dtype=torch.float32, device=self.device)
            for p, off in zip(seg.params, seg.offsets):
                ent = state_dict["state"][index[id(p)]]
                n = p.numel()
                for k in fulls:
                    fulls[k][off:off + n].copy_(ent[k].reshape(-1).to(self.device, fulls[k].dtype))
                pfull[off:off + n].copy_(ent["param"].reshape(-1).to(self.device, torch.float32))
            if seg.scales is not None:
                for k in ("exp_avg", "exp_avg_sq", "param"):
                    seg.store_scaled(k, seg.shard_view(fulls[k]).contiguous().view(-1))
            else:
                seg.exp_avg.view(seg.n_buckets, seg.shard_elems).copy_(seg.shard_view(fulls["exp_avg"]))
                seg.exp_avg_sq.view(seg.n_buckets, seg.shard_elems).copy_(seg.shard_view(fulls["exp_avg_sq"]))
                if seg.master is not None:
                    seg.master.view(seg.n_buckets, seg.shard_elems).copy_(seg.shard_view(fulls["param"]))
            if seg.remainders is not None:
                # split the fp32 master with the SAME convention as the step kernel: signed low half, high half = (bits - lo) >> 16
                bits = pfull.view(torch.int32)
                lo = ((bits & 0xFFFF) ^ 0x8000) - 0x8000
                seg.remainders.copy_(seg.shard_view(lo).contiguous().view(-1).to(torch.int16))
                seg.param_buf.view(torch.int16).copy_(((bits - lo) >> 16).to(torch.int16))
            elif not seg.param_dtype.is_floating_point:
                _pack_msb(pfull.to(seg.dtype), seg.param_buf)
            else:
                seg.param_buf.copy_(pfull.to(seg.param_dtype))
            for p in seg.params:
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
