"""Run-time switches of the scheduler (reference torchsched/config.py:9-77: env-driven, and patchable through
``torch.utils._config_module``: ``config.patch(num_streams=2)`` works as a context manager / decorator)."""
import os
import re
import sys

# print every plan and generated program
debug = os.environ.get("TORCH_SCHED_DEBUG", "0") == "1"

# rewrite passes applied to the captured graph before it is scheduled (passes/pre_grad_passes.py)
enable_pre_grad_pass = True
pre_grad_pass_options: list = ["fused_layer_norm"]

# side streams; the critical path stays on the caller's stream, everything else is dealt round-robin
num_streams = int(os.environ.get("TORCH_SCHED_NUM_STREAMS", "8"))


def _parse_ids(spec):
    """``1,2,3-5,7-10`` -> {1, 2, 3, 4, 5, 7, 8, 9, 10} (the SLURM-like syntax of the reference, config.py:23-40)."""
    out = set()
    for part in (spec or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            lo, hi = (int(x) for x in part.split("-"))
            out.update(range(lo, hi + 1))
        else:
            out.add(int(part))
    return out


# graphs (numbered in compilation order) that are left on one stream
skip_post_grad_graph_ids: set = _parse_ids(os.environ.get("TORCH_SCHED_SKIP_GRAPH_IDS"))
skip_graph_ids = skip_post_grad_graph_ids     # earlier name, same object

# give events back to a pool after their last wait and drop events nobody waits for
reuse_cuda_event: bool = os.environ.get("TORCH_SCHED_REUSE_CUDA_EVENT", "1") == "1"

# run the generated multi-stream program (inductor/) instead of interpreting the graph node by node
wrapper_codegen: bool = os.environ.get("TORCH_SCHED_CODEGEN", "0") == "1"

# trace forward AND backward with AOT autograd and schedule both graphs (backend.DecompositionsWrapper); off: only the forward graph
# dynamo captured is scheduled and backward is left to the autograd engine
aot_autograd: bool = os.environ.get("TORCH_SCHED_AOT", "0") == "1"


def _parse_dump(spec):
    """``TORCH_SCHED_DUMP_CODE='+inductor,/dir'`` -> (["torchsched", "inductor"], "/abs/dir"); without the ``+name,`` prefix only the
    torchsched program is written."""
    backends, directory = ["torchsched"], None
    m = re.fullmatch(r"(?:\+(?P<backend>\w+),)?(?P<dir>[^,]+)", (spec or "").strip())
    if m:
        if m.group("backend"):
            backends.append(m.group("backend"))
        directory = os.path.abspath(m.group("dir"))
    return backends, directory


dump_code = os.environ.get("TORCH_SCHED_DUMP_CODE", "")
dump_code_backends, dump_code_dir = _parse_dump(dump_code)

from torch.utils._config_module import install_config_module  # noqa: E402

install_config_module(sys.modules[__name__])      # adds patch(), save_config(), ...
