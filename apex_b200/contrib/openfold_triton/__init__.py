"""``apex.contrib.openfold_triton`` under its reference name: the implementation lives in :mod:`apex_b200.contrib.openfold` (there is no
Triton here; see that package's docstring). Sub-modules ``mha``, ``layer_norm`` and ``fused_adam_swa`` are importable under this name too."""
import sys

from .. import openfold as _impl
from ..openfold import *  # noqa: F401,F403
from ..openfold import __all__  # noqa: F401
from ..openfold import fused_adam_swa, layer_norm, mha  # noqa: F401

for _name in ("fused_adam_swa", "layer_norm", "mha"):
    sys.modules[f"{__name__}.{_name}"] = getattr(_impl, _name)
