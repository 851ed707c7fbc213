"""Dry run of the CUDA-only python branches on a machine without a GPU (run as a script, in its own process: it patches torch globally).

``torch.Tensor.is_cuda`` is forced to True, ``_lib.fn(name)`` returns a stub that checks the argument COUNT and TYPES against the declared
ctypes signature and does nothing, the pybind ``TensorTable`` is replaced by a python stand-in. Results are garbage (no kernel runs), but every
statement of the CUDA branches executes: imports, attribute access, argument marshalling, autograd plumbing, state_dict round trips. Used by
tests/test_cpu_abi.py; prints one line per scenario and exits non-zero if any scenario raised."""
import os
import sys
import traceback
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")

import torch  # noqa: E402

from apex_b200 import _lib  # noqa: E402

calls: list = []


class _Stub:
    def __init__(self, name):
        self.name, self.spec = name, _lib._SIGS[name]

    def __call__(self, *args):
        assert len(args) == len(self.spec), f"{self.name}: {len(args)} arguments, signature has {len(self.spec)}"
        for i, (a, c) in enumerate(zip(args, self.spec)):
            if c == "p":
                ok = a is None or (isinstance(a, int) and not isinstance(a, bool))
            elif c in "il":
                ok = isinstance(a, int)
            else:
                ok = isinstance(a, (float, int)) and not isinstance(a, bool)
            assert ok, f"{self.name}: argument {i} (code {c}) is a {type(a).__name__}"
        calls.append(self.name)
        return 0


class _Table:
    def __init__(self, lists, chunk_size=65536):
        assert chunk_size > 0 and chunk_size % 32 == 0, f"TensorTable chunk {chunk_size}"
        self.lists = [list(x) for x in lists]
        self.n, self.depth, self.chunk = len(self.lists[0]), len(self.lists), chunk_size
        self.total_chunks = sum((t.numel() + chunk_size - 1) // chunk_size for t in self.lists[0])
        self.device = self.lists[0][0].device if self.n else None
        self.dtypes = [_lib.dt(x[0]) if x else 0 for x in self.lists]
        self.uploads, self.total_numel = 1, sum(t.numel() for t in self.lists[0])

    def head(self):
        return (4096, self.n, self.depth, self.total_chunks, self.chunk)

    def track_grads(self, a, b):
        pass

    def track_universe(self, p, m):
        pass

    def refresh_grads(self):
        return True

    def set_slot(self, d, t):
        pass

    def slot(self, d):
        return self.lists[d]


_lib.fn = lambda name: _Stub(name)
_lib.available = lambda: True
_lib.stream_ptr = lambda device=None: 0
torch.Tensor.is_cuda = property(lambda self: True)

from apex_b200.ops import amp_C  # noqa: E402
import apex_b200.optimizers._base as _base  # noqa: E402

amp_C.TensorTable = _Table
_base.TensorTable = _Table

failures = 0


class _Dev(str):
    """Passes for a device wherever a string does ("cpu": allocations succeed) while reporting ``type == "cuda"`` to branch selection."""
    type, index = "cuda", 0


def attempt(label, fn, expect=()):
    global failures
    n = len(calls)
    try:
        fn()
        made = sorted(set(calls[n:]))
        missing = [k for k in expect if k not in made]
        if missing:
            failures += 1
            print("FAIL", label, "did not reach", missing)
        else:
            print("ok  ", label, made)
    except Exception as e:  # noqa: BLE001
        failures += 1
        tb = traceback.extract_tb(e.__traceback__)[-1]
        print("FAIL", label, f"{type(e).__name__}: {str(e)[:160]} @ {os.path.basename(tb.filename)}:{tb.lineno}")


def main():
    import importlib

    from apex_b200 import ext_compat
    from apex_b200 import optimizers as O
    from apex_b200.contrib.clip_grad import clip_grad_norm_
    from apex_b200.normalization import FusedLayerNorm, FusedRMSNorm
    from apex_b200.normalization import custom_ops as C

    x = torch.randn(4, 8, 16, requires_grad=True)
    attempt("FusedLayerNorm", lambda: FusedLayerNorm(16)(x).sum().backward(), ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])
    attempt("FusedRMSNorm", lambda: FusedRMSNorm(16)(x).sum().backward(), ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])
    attempt("norm custom ops", lambda: C.norm(x, torch.ones(16, requires_grad=True), torch.zeros(16, requires_grad=True), (16,), 1e-5).sum().backward(),
            ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])
    G = importlib.import_module("apex_b200.contrib.group_norm.group_norm")
    img = torch.randn(2, 8, 4, 4).contiguous(memory_format=torch.channels_last).requires_grad_()
    ones, zeros = torch.ones(8), torch.zeros(8)
    attempt("GroupNorm module", lambda: G.GroupNorm(4, 8, act="silu")(img).sum().backward(), ["ab_group_norm_small"])
    attempt("group_norm fprop / bprop", lambda: G.group_norm_nhwc_bprop(torch.ones_like(img), G.group_norm_nhwc_fprop(img.detach(), 4, ones, zeros, 1e-5, "silu")[1],
                                                                       img.detach(), 4, ones, zeros, 1e-5, "silu"), ["ab_group_norm_small"])
    attempt("group_norm custom ops", lambda: G.group_norm_nhwc_fprop_op(img, 4, torch.ones(8, requires_grad=True), torch.zeros(8, requires_grad=True), 1e-5,
                                                                       "silu")[0].sum().backward(), ["ab_group_norm_small"])
    m = ext_compat.extension_modules()
    ln, w16, b16 = m["fused_layer_norm_cuda"], torch.ones(16), torch.zeros(16)
    attempt("fused_layer_norm_cuda", lambda: ln.backward_affine(torch.randn(4, 16), *ln.forward_affine(torch.randn(4, 16), (16,), w16, b16, 1e-5)[1:], torch.randn(4, 16),
                                                               (16,), w16, b16, 1e-5), ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])
    attempt("fused_layer_norm_cuda rms", lambda: ln.rms_backward_affine(torch.randn(4, 16), ln.rms_forward_affine(torch.randn(4, 16), (16,), w16, 1e-5)[1], torch.randn(4, 16),
                                                                       (16,), w16, 1e-5), ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])
    sm = m["scaled_masked_softmax_cuda"]
    attempt("scaled_masked_softmax_cuda", lambda: sm.backward(torch.randn(2, 2, 4, 4), sm.forward(torch.randn(2, 2, 4, 4), torch.zeros(2, 1, 4, 4, dtype=torch.bool), 1.0), 1.0),
            ["ab_softmax_fwd", "ab_softmax_bwd"])
    attempt("scaled_upper_triang_masked_softmax_cuda", lambda: m["scaled_upper_triang_masked_softmax_cuda"].forward(torch.randn(3, 4, 4), 1.0), ["ab_softmax_fwd"])
    xe = m["xentropy_cuda"]
    attempt("xentropy_cuda", lambda: xe.backward(torch.ones(5), torch.randn(5, 7), xe.forward(torch.randn(5, 7), torch.randint(0, 7, (5,)), 0.1, True)[1],
                                                 torch.randint(0, 7, (5,)), 0.1), ["ab_xentropy_fwd", "ab_xentropy_bwd"])
    rope = m["fused_rotary_positional_embedding"]
    attempt("fused_rotary_positional_embedding", lambda: (rope.forward(torch.randn(6, 2, 2, 8), torch.randn(6, 1, 1, 8)),
                                                          rope.forward_2d(torch.randn(2, 3, 4, 2, 8), torch.randn(1, 3, 1, 4), torch.randn(1, 3, 1, 4), torch.randn(1, 4, 1, 4),
                                                                          torch.randn(1, 4, 1, 4)),
                                                          rope.forward_thd(torch.randn(12, 2, 8), torch.tensor([0, 5, 12], dtype=torch.int32), torch.randn(8, 1, 1, 8))), ["ab_rope"])
    fl = m["focal_loss_cuda"]
    attempt("focal_loss_cuda", lambda: fl.backward(torch.tensor(1.0), fl.forward(torch.randn(4, 30, 10), torch.randint(-2, 10, (4, 30)), torch.tensor([11.0]), 8, 0.25, 2.0,
                                                                               0.0)[1], torch.tensor([11.0])), ["ab_focal_loss_fwd", "ab_focal_loss_bwd"])
    im = m["fused_index_mul_2d"]
    attempt("fused_index_mul_2d", lambda: (im.float_forward(torch.empty(20, 5), torch.randn(7, 5), torch.randn(20, 5), torch.randint(0, 7, (20,))),
                                           im.float_backward(torch.zeros(7, 5), torch.empty(20, 5), torch.randn(20, 5), torch.randn(7, 5), torch.randn(20, 5),
                                                             torch.randint(0, 7, (20,)))), ["ab_index_mul_2d_fwd", "ab_index_mul_2d_bwd"])
    fast = m["fast_layer_norm"]
    attempt("fast_layer_norm", lambda: fast.ln_bwd(torch.randn(4, 16), torch.randn(4, 16), *fast.ln_fwd(torch.randn(4, 16), w16, b16, 1e-5)[1:], w16),
            ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])

    # ---- GEMM front-ends (bf16 so the tcgen05 path is chosen) ---------------------------------------------------------------------
    from apex_b200 import fused_dense as FD
    from apex_b200.mlp import MLP
    from apex_b200.ops import gemm as GM

    bf = torch.bfloat16
    xb = torch.randn(4, 6, 32, dtype=bf, requires_grad=True)
    attempt("FusedDense", lambda: FD.FusedDense(32, 16).to(bf)(xb).sum().backward(), ["ab_gemm_bf16", "ab_colsum"])
    attempt("FusedDense no bias", lambda: FD.FusedDense(32, 16, bias=False).to(bf)(xb).sum().backward(), ["ab_gemm_bf16"])
    attempt("FusedDenseGeluDense", lambda: FD.FusedDenseGeluDense(32, 64, 16).to(bf)(xb).sum().backward(), ["ab_gemm_bf16", "ab_colsum"])
    for act in ("none", "relu", "sigmoid"):
        attempt(f"MLP {act}", lambda act=act: MLP([32, 64, 16], activation=act).to(bf)(xb.reshape(-1, 32)).sum().backward(), ["ab_gemm_bf16"])
    attempt("MLP no bias", lambda: MLP([32, 64, 16], bias=False).to(bf)(xb.reshape(-1, 32)).sum().backward(), ["ab_gemm_bf16"])
    if hasattr(FD, "fused_dense_fp8_function"):
        w1, b1 = torch.randn(64, 32, dtype=bf, requires_grad=True), torch.zeros(64, dtype=bf, requires_grad=True)
        w2, b2 = torch.randn(16, 64, dtype=bf, requires_grad=True), torch.zeros(16, dtype=bf, requires_grad=True)
        attempt("fused_dense_fp8_function", lambda: FD.fused_dense_fp8_function(xb, w1, b1).sum().backward(), ["ab_gemm_fp8", "ab_fp8_quantize"])
        attempt("fused_dense_gelu_dense_fp8_function", lambda: FD.fused_dense_gelu_dense_fp8_function(xb, w1, b1, w2, b2).sum().backward(),
                ["ab_gemm_fp8", "ab_fp8_quantize"])
    main_grad = torch.zeros(16, 32)
    attempt("linear_wgrad accumulate fp32", lambda: GM.linear_wgrad(torch.randn(24, 16, dtype=bf), torch.randn(24, 32, dtype=bf), accum_into=main_grad), ["ab_gemm_bf16"])
    wg = m["fused_weight_gradient_mlp_cuda"]
    attempt("fused_weight_gradient_mlp_cuda", lambda: (wg.wgrad_gemm_accum_fp32(torch.randn(24, 32, dtype=bf), torch.randn(24, 16, dtype=bf), main_grad),
                                                       wg.wgrad_gemm_accum_fp16(torch.randn(24, 32, dtype=bf), torch.randn(24, 16, dtype=bf), main_grad.to(bf))), ["ab_gemm_bf16"])
    fdc = m["fused_dense_cuda"]
    attempt("fused_dense_cuda", lambda: (fdc.linear_bias_forward(torch.randn(8, 32, dtype=bf), torch.randn(16, 32, dtype=bf), torch.zeros(16, dtype=bf)),
                                         fdc.linear_bias_backward(torch.randn(8, 32, dtype=bf), torch.randn(16, 32, dtype=bf), torch.randn(8, 16, dtype=bf))), ["ab_gemm_bf16"])

    # ---- softmax / cross entropy / transducer modules -------------------------------------------------------------------------------
    from apex_b200.contrib.transducer import TransducerJoint, TransducerLoss
    from apex_b200.contrib.xentropy import SoftmaxCrossEntropyLoss
    from apex_b200.transformer.functional.fused_softmax import AttnMaskType, FusedScaleMaskSoftmax

    sc = torch.randn(2, 4, 8, 8, dtype=bf, requires_grad=True)
    attempt("FusedScaleMaskSoftmax causal", lambda: FusedScaleMaskSoftmax(False, True, AttnMaskType.causal, True, None, False, None)(sc, None).sum().backward(),
            ["ab_softmax_fwd", "ab_softmax_bwd"])
    attempt("FusedScaleMaskSoftmax padding", lambda: FusedScaleMaskSoftmax(False, True, AttnMaskType.padding, True, None, True, 2.0)(
        sc, torch.zeros(2, 1, 8, 8, dtype=torch.bool)).sum().backward(), ["ab_softmax_fwd", "ab_softmax_bwd"])
    lg = torch.randn(6, 11, requires_grad=True)
    attempt("SoftmaxCrossEntropyLoss", lambda: SoftmaxCrossEntropyLoss.apply(lg, torch.randint(0, 11, (6,)), 0.1, 0, True).sum().backward(),
            ["ab_xentropy_fwd", "ab_xentropy_bwd"])
    f, g = torch.randn(2, 5, 8, requires_grad=True), torch.randn(2, 3, 8, requires_grad=True)
    fl_, gl_ = torch.tensor([5, 4], dtype=torch.int32), torch.tensor([3, 2], dtype=torch.int32)
    attempt("TransducerJoint", lambda: TransducerJoint()(f, g, fl_, gl_).sum().backward(), ["ab_transducer_joint_fwd", "ab_transducer_joint_bwd"])
    attempt("TransducerJoint relu + dropout", lambda: TransducerJoint(relu=True, dropout=True, dropout_prob=0.1)(f, g, fl_, gl_).sum().backward(),
            ["ab_transducer_joint_fwd", "ab_transducer_joint_bwd"])
    lx = torch.randn(2, 5, 4, 7, requires_grad=True)
    attempt("TransducerLoss", lambda: TransducerLoss()(lx, torch.randint(1, 7, (2, 3)), fl_, gl_, 0).sum().backward(), ["ab_transducer_loss_fwd", "ab_transducer_loss_bwd"])

    def drive(make, late=False):
        ps = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(4))]
        opt = make(ps)
        for it in range(3):
            ps[0].grad = torch.randn(5, 3)
            if not late or it > 0:
                ps[1].grad = torch.randn(4)
            opt.step()
        opt.load_state_dict(opt.state_dict())
        ps[0].grad, ps[1].grad = torch.randn(5, 3), torch.randn(4)
        opt.step()

    attempt("FusedAdam", lambda: drive(lambda p: O.FusedAdam(p, lr=1e-2)), ["ab_mt_adam"])
    attempt("FusedAdam capturable + master", lambda: drive(lambda p: O.FusedAdam(p, lr=1e-2, capturable=True, master_weights=True)), ["ab_mt_adam"])
    attempt("FusedSGD", lambda: drive(lambda p: O.FusedSGD(p, lr=1e-2, momentum=0.9)), ["ab_mt_sgd"])
    attempt("FusedSGD late gradient", lambda: drive(lambda p: O.FusedSGD(p, lr=1e-2, momentum=0.9), late=True), ["ab_mt_sgd"])
    attempt("FusedLAMB", lambda: drive(lambda p: O.FusedLAMB(p, lr=1e-2)), ["ab_mt_lamb_stage1", "ab_mt_lamb_stage2"])
    attempt("FusedMixedPrecisionLamb", lambda: drive(lambda p: O.FusedMixedPrecisionLamb(p, lr=1e-2)), ["ab_mt_lamb_stage1", "ab_mt_lamb_stage2"])
    attempt("FusedNovoGrad", lambda: drive(lambda p: O.FusedNovoGrad(p, lr=1e-2)), ["ab_mt_novograd"])
    attempt("FusedAdagrad", lambda: drive(lambda p: O.FusedAdagrad(p, lr=1e-2)), ["ab_mt_adagrad"])

    def clip():
        p = torch.nn.Parameter(torch.randn(5, 3))
        p.grad = torch.randn(5, 3)
        clip_grad_norm_([p], 1.0)

    attempt("clip_grad_norm_", clip, ["ab_mt_norm"])

    def distopt():
        p, mm, v, g = torch.randn(100), torch.zeros(100), torch.zeros(100), torch.randn(100)
        noop, one = torch.zeros(1, dtype=torch.int32), torch.ones(1)
        d = m["distributed_adam_cuda"]
        d.multi_tensor_fused_adam(65536, noop, [[p], [mm], [v], [g], [p]], one, 1e-2, 0.9, 0.99, 1e-8, 1, 1, 1, 0.1)
        d.multi_tensor_fused_adam_capturable(65536, noop, [[p], [mm], [v], [g], [p]], one, torch.tensor([1e-2]), 0.9, 0.99, 1e-8, torch.ones(1, dtype=torch.int32), 1, 1, 0.1)
        pb, rem = torch.randn(100).bfloat16(), torch.zeros(100, dtype=torch.int16)
        d.multi_tensor_fused_adam_with_param_remainders(65536, noop, [[pb], [rem], [mm], [v], [g], [pb]], one, 1e-2, 0.9, 0.99, 1e-8, 1, 1, 1, 0.1)

    attempt("distributed_adam_cuda", distopt, ["ab_mt_dist_adam", "ab_mt_dist_adam_remainders"])
    # ---- remaining single-process entry points ---------------------------------------------------------------------------------------
    big = torch.randn(1, 8, 512, 512).contiguous(memory_format=torch.channels_last).requires_grad_()
    attempt("GroupNorm large slab", lambda: G.GroupNorm(4, 8)(big).sum().backward(), ["ab_group_norm"])
    from apex_b200.parallel import SyncBatchNorm
    from apex_b200.parallel import sync_batchnorm as SB

    _get = SB._GroupState.get.__func__
    SB._GroupState.get = classmethod(lambda cls, group, device, channels: _get(cls, group, _Dev("cpu"), channels))

    xi = torch.randn(4, 6, 5, 5, requires_grad=True)
    attempt("SyncBatchNorm training", lambda: SyncBatchNorm(6)(xi).sum().backward(), ["ab_syncbn"])
    attempt("SyncBatchNorm channels-last + relu + residual", lambda: SyncBatchNorm(6, channel_last=True, fuse_relu=True)(
        xi.permute(0, 2, 3, 1).contiguous(), torch.randn(4, 5, 5, 6)).sum().backward(), ["ab_syncbn"])
    attempt("SyncBatchNorm eval (running statistics: F.batch_norm)", lambda: SyncBatchNorm(6).eval()(xi).sum().backward())
    noop = torch.zeros(1, dtype=torch.int32)
    a, b, c = [torch.randn(40), torch.randn(7)], [torch.randn(40), torch.randn(7)], [torch.empty(40), torch.empty(7)]
    attempt("amp_C scale / axpby / cast", lambda: (amp_C.multi_tensor_scale(65536, noop, [a, c], 0.5), amp_C.multi_tensor_axpby(65536, noop, [a, b, c], 1.0, 2.0, -1),
                                                  amp_C.multi_tensor_cast(65536, noop, [a, [t.bfloat16() for t in c]])), ["ab_mt_scale", "ab_mt_axpby", "ab_mt_cast"])
    attempt("amp_C norms", lambda: (amp_C.multi_tensor_l2norm(65536, noop, [a], True), amp_C.multi_tensor_l2norm_mp(65536, noop, [a], False),
                                   amp_C.multi_tensor_unscale_l2norm(65536, noop, [a], torch.ones(1), True), amp_C.multi_tensor_l2norm_scale(65536, noop, [a, c], 0.5, True),
                                   amp_C.multi_tensor_norm_out(65536, noop, [a, [torch.zeros(2)]], torch.zeros(2), 0.5, 0.5, 0)), ["ab_mt_norm", "ab_mt_l2norm_scale"])
    attempt("amp_C update_scale_hysteresis", lambda: amp_C.update_scale_hysteresis(torch.ones(1), torch.zeros(1, dtype=torch.int32), torch.ones(1, dtype=torch.int32),
                                                                                  torch.zeros(1), 2.0, 0.5, 100, 2), ["ab_update_scale_hysteresis"])

    import numpy as np

    from apex_b200.contrib.sparsity import permutation_search as PS

    wm = torch.randn(16, 32)
    attempt("permutation scoring kernels", lambda: (PS.sum_after_2_to_4(wm), PS.sum_after_2_to_4(wm, torch.stack([torch.randperm(32) for _ in range(3)])),
                                                   PS._score_groups(wm, PS._all_groups(8, 2, "cpu"), np.stack([np.random.permutation(8) for _ in range(5)]).astype(np.uint8))),
            ["ab_perm_eval", "ab_stripe_search"])

    # ---- module / functional forms of the small fused ops ------------------------------------------------------------------------------
    from apex_b200.contrib.layer_norm import FastLayerNorm
    from apex_b200.normalization import MixedFusedLayerNorm, MixedFusedRMSNorm
    from apex_b200.transformer.functional import fused_rope as RP

    FLm = importlib.import_module("apex_b200.contrib.focal_loss.focal_loss")
    IMm = importlib.import_module("apex_b200.contrib.index_mul_2d.index_mul_2d")
    xh = torch.randn(4, 8, 16, dtype=torch.bfloat16, requires_grad=True)
    attempt("FastLayerNorm", lambda: FastLayerNorm(16)(x).sum().backward(), ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])
    attempt("MixedFusedLayerNorm bf16 input", lambda: MixedFusedLayerNorm(16)(xh).sum().backward(), ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])
    attempt("MixedFusedRMSNorm memory efficient", lambda: MixedFusedRMSNorm(16, memory_efficient=True)(xh).sum().backward(), ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])
    attempt("FusedLayerNorm no affine + memory efficient", lambda: (FusedLayerNorm(16, elementwise_affine=False)(x).sum().backward(),
                                                                   FusedLayerNorm(16, memory_efficient=True)(x).sum().backward()), ["ab_layer_norm_fwd", "ab_layer_norm_bwd"])
    co = torch.randn(4, 30, 10, requires_grad=True)
    attempt("focal_loss", lambda: FLm.focal_loss(co, torch.randint(-2, 10, (4, 30)), torch.tensor([11.0]), 8, 0.25, 2.0, 0.1).backward(),
            ["ab_focal_loss_fwd", "ab_focal_loss_bwd"])
    i1, i2 = torch.randn(7, 5, requires_grad=True), torch.randn(20, 5, requires_grad=True)
    attempt("index_mul_2d (+ double backward)", lambda: torch.autograd.grad(torch.autograd.grad(IMm.index_mul_2d(i1, i2, torch.randint(0, 7, (20,))).sum(), i1,
                                                                                               create_graph=True)[0].sum(), i2), ["ab_index_mul_2d_fwd", "ab_index_mul_2d_bwd"])
    tq = torch.randn(6, 2, 2, 8, requires_grad=True)
    fr = torch.randn(6, 1, 1, 8)
    attempt("RoPE sbhd / cached / transposed", lambda: (RP.fused_apply_rotary_pos_emb(tq, fr).sum().backward(),
                                                       RP.fused_apply_rotary_pos_emb(tq, fr, transpose_output_memory=True).sum().backward(),
                                                       RP.fused_apply_rotary_pos_emb_cached(tq, fr.cos(), fr.sin()).sum().backward()), ["ab_rope"])
    tt = torch.randn(12, 2, 8, requires_grad=True)
    attempt("RoPE thd", lambda: RP.fused_apply_rotary_pos_emb_thd(tt, torch.tensor([0, 5, 12], dtype=torch.int32), torch.randn(8, 1, 1, 8)).sum().backward(), ["ab_rope"])
    t2 = torch.randn(2, 12, 2, 8, requires_grad=True)
    attempt("RoPE 2d", lambda: RP.fused_apply_rotary_pos_emb_2d(t2, 3, 4, torch.randn(1, 3, 1, 4), torch.randn(1, 3, 1, 4), torch.randn(1, 4, 1, 4),
                                                               torch.randn(1, 4, 1, 4)).sum().backward(), ["ab_rope"])

    # ---- DistributedFusedAdam: one rank, device object that reports type "cuda" but allocates on the CPU --------------------------
    import torch.distributed as dist

    from apex_b200.contrib.optimizers.distributed_fused_adam import DistributedFusedAdam

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29613")
    dist.init_process_group("gloo", rank=0, world_size=1)

    def dfa(fused, pdtype=torch.float32, **kw):
        ps = [torch.nn.Parameter(torch.randn(33, 7, dtype=pdtype)), torch.nn.Parameter(torch.randn(50, dtype=pdtype))]
        opt = DistributedFusedAdam(ps, lr=1e-2, device="cpu", **kw)
        opt.device, opt.fused_collectives = _Dev("cpu"), fused
        for _ in range(2):
            for q in ps:
                q.grad = torch.randn_like(q)
            opt.clip_grad_norm(1.0)
            opt.step()
            opt.zero_grad()
        sd = opt.state_dict()
        opt.load_state_dict(sd)
        for q in ps:
            q.grad = torch.randn_like(q)
        opt.step()

    attempt("DistributedFusedAdam one-kernel path", lambda: dfa(True), ["ab_dist_adam_step"])
    attempt("DistributedFusedAdam one-kernel path bf16", lambda: dfa(True, torch.bfloat16), ["ab_dist_adam_step"])
    attempt("DistributedFusedAdam capturable", lambda: dfa(True, capturable=True), ["ab_dist_adam_step"])
    attempt("DistributedFusedAdam generic path", lambda: dfa(False), ["ab_mt_dist_adam"])
    attempt("DistributedFusedAdam remainders", lambda: dfa(False, torch.bfloat16, store_params=False, store_param_remainders=True), ["ab_mt_dist_adam_remainders"])
    attempt("DistributedFusedAdam scaled states", lambda: dfa(False, torch.bfloat16, dtype=torch.bfloat16, with_scaled_states=True), ["ab_mt_dist_adam"])
    attempt("DistributedFusedAdam integer param sync", lambda: dfa(False, param_sync_dtype=torch.uint8), ["ab_mt_dist_adam"])
    dist.destroy_process_group()
    return failures


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
