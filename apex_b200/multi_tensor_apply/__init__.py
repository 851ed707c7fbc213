"""``multi_tensor_applier`` — same calling convention as the reference (apex/multi_tensor_apply/multi_tensor_apply.py:24-27):
``multi_tensor_applier(op, noop_flag, tensor_lists, *args)`` with ``op`` one of ``apex_b200.ops.amp_C.*``."""
from .. import _lib


class MultiTensorApply:
    available = True
    warned = False

    def __init__(self, chunk_size):
        self.chunk_size = chunk_size

    def check_avail(self):
        """Raises when CUDA tensors could not be served (reference :15-22). CPU-only machines use the PyTorch oracles, so this only
        fails on a GPU machine whose native library did not load."""
        import torch

        if torch.cuda.is_available():
            _lib.require()

    def __call__(self, op, noop_flag_buffer, tensor_lists, *args):
        return op(self.chunk_size, noop_flag_buffer, tensor_lists, *args)


multi_tensor_applier = MultiTensorApply(2048 * 32)
__all__ = ["MultiTensorApply", "multi_tensor_applier"]
