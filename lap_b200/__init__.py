"""lap_b200 — B200-native (sm_100a) engine for the LAP-3B hot path of lihzha/lap.

Importing the package never touches the GPU; the kernel library is loaded (and, if stale, rebuilt in-tree) on first
use.  There is no CPU fallback.
"""
from .config import LAPConfig, TrainConfig, get_config  # noqa: F401
from .observation import CoTObservation, Observation  # noqa: F401

__all__ = ["LAPConfig", "TrainConfig", "get_config", "Observation", "CoTObservation"]
