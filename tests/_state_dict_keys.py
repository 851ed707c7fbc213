"""Helper of tests/test_cpu_modules.py::test_state_dict_keys_match_the_reference_constructors: prints {module case: {state_dict key: shape}} as JSON
for either the reference package (argv[1] == "ref": /root/reference on sys.path, compiled extensions stubbed - only constructors run) or this one."""
import sys, json, types, warnings
warnings.simplefilter("ignore")
which = sys.argv[1]
if which == "ref":
    sys.path.insert(0, "/root/reference")
    # compiled extensions are not importable here: stub them so module-level imports succeed (only constructors run)
    import importlib.abc, importlib.machinery
    class Stub(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        names = {"fused_layer_norm_cuda","fast_layer_norm","fused_dense_cuda","mlp_cuda","group_norm_cuda","group_norm_v2_cuda","bnp","fast_multihead_attn","fmhalib","transducer_joint_cuda","transducer_loss_cuda","focal_loss_cuda","fast_bottleneck","nccl_p2p_cuda","peer_memory_cuda","cudnn_gbn_lib","xentropy_cuda","fused_conv_bias_relu","amp_C","syncbn","fused_index_mul_2d","apex_C"}
        def find_spec(self, name, path=None, target=None):
            return importlib.machinery.ModuleSpec(name, self) if name in self.names else None
        def create_module(self, spec): return types.ModuleType(spec.name)
        def exec_module(self, m): pass
    sys.meta_path.insert(0, Stub())
    pkg = "apex"
else:
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    pkg = "apex_b200"
import importlib, torch
def cls(path):
    mod, name = path.rsplit(".", 1)
    return getattr(importlib.import_module(pkg + "." + mod), name)
cfg = types.SimpleNamespace(attention_probs_dropout_prob=0.1, num_attention_heads=4, hidden_size=64)
cases = {
 "normalization.FusedLayerNorm": ((32,), {}), "normalization.FusedRMSNorm": ((32,), {}), "normalization.MixedFusedLayerNorm": ((32,), {}),
 "normalization.MixedFusedRMSNorm": ((32,), {}), "normalization.FusedLayerNorm#noaffine": ((32,), {"elementwise_affine": False}),
 "fused_dense.FusedDense": ((16, 32), {}), "fused_dense.FusedDense#nobias": ((16, 32), {"bias": False}), "fused_dense.FusedDenseGeluDense": ((16, 32, 8), {}),
 "mlp.MLP": (([16, 32, 8],), {}), "mlp.MLP#nobias": (([16, 32, 8],), {"bias": False}),
 "contrib.multihead_attn.SelfMultiheadAttn": ((32, 4), {"bias": True}), "contrib.multihead_attn.SelfMultiheadAttn#norm": ((32, 4), {"include_norm_add": True}),
 "contrib.multihead_attn.SelfMultiheadAttn#sep": ((32, 4), {"bias": True, "separate_qkv_params": True}),
 "contrib.multihead_attn.SelfMultiheadAttn#default_norm": ((32, 4), {"include_norm_add": True, "impl": "default"}),
 "contrib.multihead_attn.EncdecMultiheadAttn": ((32, 4), {"bias": True}), "contrib.multihead_attn.EncdecMultiheadAttn#norm": ((32, 4), {"include_norm_add": True}),
 "contrib.group_norm.GroupNorm": ((4, 32), {}), "contrib.layer_norm.FastLayerNorm": ((1024,), {}),
 "contrib.groupbn.BatchNorm2d_NHWC": ((16,), {}), "contrib.cudnn_gbn.GroupBatchNorm2d": ((16, 1), {}),
 "contrib.bottleneck.Bottleneck": ((16, 8, 32), {}), "contrib.fmha.FMHA": ((cfg,), {}),
 "contrib.transducer.TransducerJoint": ((), {}), "contrib.transducer.TransducerLoss": ((), {}),
 "contrib.xentropy.SoftmaxCrossEntropyLoss": None,
}
out = {}
for key, spec in cases.items():
    if spec is None: continue
    path = key.split("#")[0]
    try:
        m = cls(path)(*spec[0], **spec[1])
        out[key] = {k: list(v.shape) for k, v in m.state_dict().items()}
    except Exception as e:
        out[key] = "ERR " + repr(e)[:150]
print(json.dumps(out))
