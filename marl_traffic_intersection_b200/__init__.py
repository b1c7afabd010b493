"""Importable alias of the package directory ``marl-traffic-intersection_b200/`` (a hyphen cannot be
imported).  All code lives there; this file only points ``__path__`` at it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "marl-traffic-intersection_b200")
__path__.insert(0, _real)

from ._lib import LIB_PATH, IsxError, load_library  # noqa: E402,F401
from .batched import BatchedIntersectionEnv  # noqa: E402,F401
from .env import DEFAULT_REWARD_CONFIG, IntersectionEnv  # noqa: E402,F401
from .utils import DEFAULT_ROUTE_MAPPING_2LANES, DEFAULT_ROUTE_MAPPING_3LANES  # noqa: E402,F401
