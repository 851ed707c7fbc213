from .mlp import MLP, mlp_function

__all__ = ["MLP", "mlp_function"]
