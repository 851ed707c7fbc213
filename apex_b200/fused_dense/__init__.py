from .fused_dense import (DenseNoBiasFunc, FusedDense, FusedDenseFP8Func, FusedDenseGeluDenseFP8Func, fused_dense_gelu_dense_fp8_function, FusedDenseFunc, FusedDenseGeluDense, FusedDenseGeluDenseFunc,
                          fused_dense_fp8_function, fused_dense_function, fused_dense_gelu_dense_function)

__all__ = ["FusedDense", "FusedDenseGeluDense", "FusedDenseFunc", "DenseNoBiasFunc", "FusedDenseGeluDenseFunc", "fused_dense_function",
           "fused_dense_gelu_dense_function", "FusedDenseFP8Func", "fused_dense_fp8_function", "FusedDenseGeluDenseFP8Func",
           "fused_dense_gelu_dense_fp8_function"]
