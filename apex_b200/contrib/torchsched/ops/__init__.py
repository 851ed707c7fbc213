from .layer_norm import layer_norm  # noqa: F401
