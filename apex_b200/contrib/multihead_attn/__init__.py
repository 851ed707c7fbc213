from .multihead_attn import EncdecMultiheadAttn, SelfMultiheadAttn, fast_mask_softmax_dropout_func

__all__ = ["SelfMultiheadAttn", "EncdecMultiheadAttn", "fast_mask_softmax_dropout_func"]
