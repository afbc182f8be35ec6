"""townlet_b200 — B200-native per-tick agent hot path of Townlet.

Observation build (hybrid / full / compact map + feature vector + social
snippet), need decay and reward, batched over independent towns, as sm_100a
CUDA kernels behind a C ABI (``include/townlet_b200.h``).  See DESIGN.md.
"""

from . import capi
from .params import DynamicsSpec, ObsSpec, RewardSpec
from .soa import BatchLayout, TownBatch

__all__ = ["BatchLayout", "DynamicsSpec", "ObsSpec", "RewardSpec", "TownBatch", "capi"]
__version__ = "0.1.0"
