from .fused_dense import (DenseNoBiasFunc, FusedDense, FusedDenseFunc, FusedDenseGeluDense, FusedDenseGeluDenseFunc, fused_dense_function,
                          fused_dense_gelu_dense_function)

__all__ = ["FusedDense", "FusedDenseGeluDense", "FusedDenseFunc", "DenseNoBiasFunc", "FusedDenseGeluDenseFunc", "fused_dense_function",
           "fused_dense_gelu_dense_function"]
