"""Program generation for scheduled graphs (import-path parity with apex/contrib/torchsched/inductor/__init__.py:1-5). The name is the
reference's; nothing here depends on Inductor or Triton: the programs call this library's kernels and ATen directly."""
from .graph import lower_graph, patch_graph_lowering

__all__ = ["patch_graph_lowering", "lower_graph"]
