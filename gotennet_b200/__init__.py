"""gotennet_b200 — B200-native (sm_100a) implementation of GotenNet's interaction path.

Drop-in for the reference package root (reference gotennet/__init__.py:5-10):
`EQFF, GATA, GotenNet, GotenNetWrapper`.  All arithmetic runs in hand-written CUDA
kernels behind the C ABI in include/gotennet_b200.h; there is no CPU, Triton or
eager-PyTorch fallback — calling a forward without the built library or on a CPU
tensor raises `GotenError`.
"""
__version__ = "0.1.0"

from ._lib import GotenError  # noqa: F401
from .gotennet import EQFF, GATA, GotenNet, GotenNetWrapper  # noqa: F401
from .layers import CosineCutoff, Dense, Distance, ExpNormalSmearing, MLP, TensorInit  # noqa: F401
from .outputs import (Atomwise, Dipole, ElectronicSpatialExtentV2, GatedEquivariantBlock, ScaleShift,  # noqa: F401
                      SchnetMLP, shifted_softplus)
from .optim import FusedAdamW  # noqa: F401
from .parallel import FlatGradBuffer, shard_bounds, take_shard  # noqa: F401
from .data import MoleculeBatch, collate  # noqa: F401
from .training import energy_and_forces, force_matching_backward  # noqa: F401
