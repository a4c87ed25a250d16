"""Galaxy components: models of the joint density p(t, v, x, z).

Same public names as the reference package (/root/reference/popkinmocks/components/__init__.py:1-6).
"""
from .base import Component  # noqa: F401
from .parametric import ParametricComponent  # noqa: F401
from .growing_disk import GrowingDisk  # noqa: F401
from .stream import Stream  # noqa: F401
from .mixture import Mixture  # noqa: F401
from .particle import FromParticle  # noqa: F401
