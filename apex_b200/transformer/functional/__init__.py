from .fused_weight_gradient import wgrad_gemm_accum_fp16, wgrad_gemm_accum_fp32  # noqa: F401
from .fused_softmax import (AttnMaskType, FusedScaleMaskSoftmax, GenericScaledMaskedSoftmax, ScaledMaskedSoftmax, ScaledSoftmax,  # noqa: F401
                            ScaledUpperTriangMaskedSoftmax, scaled_masked_softmax, scaled_softmax, scaled_upper_triang_masked_softmax)
from .fused_rope import (fused_apply_rotary_pos_emb, fused_apply_rotary_pos_emb_2d, fused_apply_rotary_pos_emb_cached,  # noqa: F401
                         fused_apply_rotary_pos_emb_thd)
