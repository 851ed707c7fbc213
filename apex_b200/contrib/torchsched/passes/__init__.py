from .pre_grad_passes import pre_grad_custom_pass, replace_layer_norm  # noqa: F401
