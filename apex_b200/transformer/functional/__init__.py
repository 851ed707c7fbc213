from .fused_weight_gradient import wgrad_gemm_accum_fp16, wgrad_gemm_accum_fp32  # noqa: F401
