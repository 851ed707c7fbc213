"""``apex.parallel.DistributedDataParallel`` / ``Reducer`` (python layer removed from the reference snapshot; API reconstructed
from tests/distributed/DDP/ddp_race_condition_test.py:39 and README.md:77-81).

Gradients are flattened into buckets of ``message_size`` elements in the order they become ready during backward; each full
bucket is all-reduced on one of ``num_allreduce_streams`` side streams while backward continues (``delay_allreduce=True`` does one
flat all-reduce at the end). ``gradient_predivide_factor`` splits the averaging around the reduction. The bucket all-reduce itself
is NCCL (gloo on CPU): it runs concurrently with backward compute, which is exactly the situation where a spinning in-kernel
collective must not be used.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def flat_dist_call(tensors, call, extra_args=None):
    """Apply a torch.distributed collective to a list of tensors through one flat buffer per dtype."""
    buckets: dict = {}
    for t in tensors:
        buckets.setdefault(t.dtype, []).append(t)
    for ts in buckets.values():
        flat = _flatten_dense_tensors(ts)
        if extra_args is not None:
            call(flat, *extra_args)
        else:
            call(flat)
        for t, synced in zip(ts, _unflatten_dense_tensors(flat, ts)):
            t.copy_(synced)


class Reducer:
    """Manual gradient averaging: call ``reducer.reduce()`` after backward (reference ``apex.parallel.Reducer``)."""

    def __init__(self, module_or_grads_list, process_group=None):
        self.process_group = process_group
        if isinstance(module_or_grads_list, torch.nn.Module):
            self.module = module_or_grads_list
            flat_dist_call([p.data for p in self.module.parameters()], dist.broadcast, (0, process_group))
        else:
            self.module = None
            self.grads = list(module_or_grads_list)

    def reduce(self):
        grads = [p.grad.data for p in self.module.parameters() if p.grad is not None] if self.module is not None else self.grads
        world = dist.get_world_size(self.process_group)
        flat_dist_call(grads, lambda t: (dist.all_reduce(t, group=self.process_group), t.div_(world)))


class DistributedDataParallel(torch.nn.Module):
    def __init__(self, module, message_size=10000000, delay_allreduce=False, shared_param=None, allreduce_trigger_params=None,
                 retain_allreduce_buffers=False, allreduce_always_fp32=False, num_allreduce_streams=1, allreduce_communicators=None,
                 gradient_average=True, gradient_predivide_factor=1.0, gradient_average_split_factor=None, prof=False,
                 process_group=None, fused_collectives=False):
        super().__init__()
        # EXPERIMENTAL: bucket all-reduce through the NVSwitch (parallel/nvls_allreduce.py) instead of NCCL; needs the experimental build
        self.fused_collectives = bool(fused_collectives)
        self._nvls = None
        if shared_param is not None:
            raise ValueError("shared_param is no longer supported as an option; use delay_allreduce=True for shared parameters")
        self.module = module
        self.process_group = process_group
        self.world_size = float(dist.get_world_size(process_group)) if dist.is_initialized() else 1.0
        self.message_size = message_size
        self.delay_allreduce = delay_allreduce
        self.retain_allreduce_buffers = retain_allreduce_buffers
        self.allreduce_always_fp32 = allreduce_always_fp32
        self.gradient_average = gradient_average
        self.gradient_predivide_factor = gradient_predivide_factor
        self.num_allreduce_streams = max(1, num_allreduce_streams)
        self.allreduce_trigger_params = set(id(p) for p in allreduce_trigger_params) if allreduce_trigger_params else None
        self._streams = None
        self._pending: list = []
        self._ready: list = []
        self._ready_elems = 0
        self._bucket_idx = 0
        self._disabled = False
        self._callback_queued = False
        self.allreduce_buffers: list = []
        params = [p for p in module.parameters()]
        if dist.is_initialized() and self.world_size > 1:
            flat_dist_call([p.data for p in params], dist.broadcast, (dist.get_global_rank(process_group, 0) if process_group else 0, process_group))
        self._params = [p for p in params if p.requires_grad]
        self._n_expected = len(self._params)
        self._n_seen = 0
        for p in self._params:
            p.register_post_accumulate_grad_hook(self._make_hook())

    # ------------------------------------------------------------------------------------------------------------------
    def _make_hook(self):
        def hook(param):
            if self._disabled or not (dist.is_initialized() and self.world_size > 1):
                return
            if not self._callback_queued:
                # parameters that take no part in this backward never fire their hook: whatever is still pending when the engine
                # finishes is reduced from an end-of-backward callback (the reference queues the same callback from its grad hooks)
                self._callback_queued = True
                torch.autograd.Variable._execution_engine.queue_callback(self._end_of_backward)
            self._n_seen += 1
            self._ready.append(param)
            self._ready_elems += param.numel()
            trigger = (self.allreduce_trigger_params is not None and id(param) in self.allreduce_trigger_params) or \
                      (self.allreduce_trigger_params is None and self._ready_elems >= self.message_size)
            last = self._n_seen == self._n_expected
            if self.delay_allreduce:
                if last:
                    self._flush(final=True)
            elif trigger or last:
                self._flush(final=last)
        return hook

    def _end_of_backward(self):
        self._callback_queued = False
        if self._ready or self._n_seen:
            self._flush(final=True)

    def _flush(self, final: bool):
        params, self._ready, self._ready_elems = self._ready, [], 0
        if params:
            self._allreduce_bucket([p.grad for p in params])
        if final:
            self._n_seen = 0
            self._bucket_idx = 0
            if self._streams is not None:
                cur = torch.cuda.current_stream()
                for s in self._streams:
                    cur.wait_stream(s)

    def _allreduce_bucket(self, grads):
        by_dtype: dict = {}
        for g in grads:
            by_dtype.setdefault(g.dtype, []).append(g)
        for dtype, gs in by_dtype.items():
            cuda = gs[0].is_cuda
            stream = None
            if cuda and not self.delay_allreduce:
                if self._streams is None:
                    self._streams = [torch.cuda.Stream() for _ in range(self.num_allreduce_streams)]
                stream = self._streams[self._bucket_idx % self.num_allreduce_streams]
                stream.wait_stream(torch.cuda.current_stream())
            self._bucket_idx += 1
            ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
            with ctx:
                flat = _flatten_dense_tensors(gs)
                if self.allreduce_always_fp32:
                    flat = flat.float()
                if self.gradient_predivide_factor != 1.0:
                    flat.mul_(1.0 / self.gradient_predivide_factor)
                if not self._allreduce_nvls(flat):
                    dist.all_reduce(flat, group=self.process_group)
                if self.gradient_average:
                    flat.mul_(self.gradient_predivide_factor / self.world_size)
                if self.retain_allreduce_buffers:
                    self.allreduce_buffers.append(flat)
                for g, synced in zip(gs, _unflatten_dense_tensors(flat, gs)):
                    g.copy_(synced)
                    if stream is not None:
                        g.record_stream(stream)

    def _allreduce_nvls(self, flat) -> bool:
        if not (self.fused_collectives and flat.is_cuda):
            return False
        from . import nvls_allreduce as NV

        if self._nvls is None:
            if not NV.available(self.process_group):
                self.fused_collectives = False
                return False
            self._nvls = NV.NvlsAllReduce(self.process_group, flat.device, max(64 << 20, flat.numel() * flat.element_size() * 2))
        if flat.numel() * flat.element_size() > self._nvls.mem.nbytes - 4096:
            return False
        self._nvls.allreduce_(flat)
        return True

    def disable_allreduce(self):
        self._disabled = True

    def enable_allreduce(self):
        self._disabled = False

    def forward(self, *inputs, **kwargs):
        self._n_seen = 0
        self._ready, self._ready_elems = [], 0
        self.allreduce_buffers = []
        return self.module(*inputs, **kwargs)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
