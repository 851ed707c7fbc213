"""Megatron-style building blocks whose kernels survive in the reference (csrc/megatron): fused weight-gradient accumulation,
scaled (masked / causal) softmax, rotary position embedding."""
from . import functional  # noqa: F401
