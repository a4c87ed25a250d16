"""popkinmocks_b200 - B200-native datacube synthesis behind the popkinmocks API.

Drop-in surface of the reference package for the `evaluate_ybar` hot path
(/root/reference/popkinmocks/__init__.py:1-3): `milesSSPs`, `ifu_cube`, `components`,
`noise`.  Arithmetic runs in libpkm_b200.so (hand-written sm_100a CUDA); there is no
CPU fallback.
"""
from .model_grids import milesSSPs  # noqa: F401
from . import ifu_cube  # noqa: F401
from . import engine  # noqa: F401
from . import components  # noqa: F401
from . import noise  # noqa: F401
