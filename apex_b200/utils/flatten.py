"""``apex_C.flatten / unflatten`` (reference csrc/flatten_unflatten.cpp:5-14: thin wrappers over torch's dense-tensor flatten)."""
from torch._utils import _flatten_dense_tensors as flatten  # noqa: F401
from torch._utils import _unflatten_dense_tensors as unflatten  # noqa: F401
