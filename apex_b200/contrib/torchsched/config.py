"""Run-time switches of the scheduler (reference torchsched/config.py:9-77: env-driven, patchable)."""
import os

debug = os.environ.get("TORCH_SCHED_DEBUG", "0") == "1"
num_streams = int(os.environ.get("TORCH_SCHED_NUM_STREAMS", "8"))
skip_graph_ids = [int(x) for x in os.environ.get("TORCH_SCHED_SKIP_GRAPH_IDS", "").split(",") if x]
reuse_cuda_event = os.environ.get("TORCH_SCHED_REUSE_CUDA_EVENT", "1") == "1"
dump_code = os.environ.get("TORCH_SCHED_DUMP_CODE", "")
