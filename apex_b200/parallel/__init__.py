"""Data-parallel building blocks: SyncBatchNorm (in-kernel cross-GPU reduction), symmetric memory over NVLink/NVSwitch,
DistributedDataParallel / Reducer (flat-bucket gradient all-reduce)."""
from .sync_batchnorm import SyncBatchNorm, convert_syncbn_model, create_syncbn_process_group  # noqa: F401
from .distributed import DistributedDataParallel, Reducer, flat_dist_call  # noqa: F401
from . import symmetric  # noqa: F401
from . import syncbn_ops  # noqa: F401  (the reference extension's decomposed entry points)
