"""``apex._autocast_utils`` (reference apex/_autocast_utils.py:1-26): cast the arguments of a custom autograd function to the
active autocast dtype, because ``torch.autograd.Function.apply`` does not take part in autocast on its own."""
from typing import Optional, Sequence

import torch

__all__ = ["_cast_if_autocast_enabled"]


def _get_autocast_dtypes() -> Sequence[torch.dtype]:
    return [torch.half, torch.bfloat16] if (torch.cuda.is_available() and torch.cuda.is_bf16_supported()) else [torch.half]


def _get_current_dtype(dtype: Optional[torch.dtype] = None) -> torch.dtype:
    if not torch.is_autocast_enabled():
        return dtype or torch.float
    return torch.get_autocast_dtype("cuda")


def _cast_if_autocast_enabled(*args):
    if not torch.is_autocast_enabled():
        return args
    dt = torch.get_autocast_dtype("cuda")
    return tuple(a.to(dt) if (torch.is_tensor(a) and a.is_floating_point() and a.is_cuda) else a for a in args)
