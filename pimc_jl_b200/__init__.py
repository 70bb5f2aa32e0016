"""pimc_jl_b200 -- B200-native (sm_100a) path integral Monte Carlo engine behind the API of oameye/PIMC.jl.

Layout: csrc/ (CUDA kernels + C ABI -> libpimc_b200.so), _lib.py (ctypes binding), engine.py (numpy-facing
handle), pimc.py (host-side mirror of the reference's Julia interface: System, run!, update and measurement functors).
"""
from . import _lib
from ._lib import PimcError, make_potential, SCHED_FAITHFUL, SCHED_SWEEP  # noqa: F401
from .engine import Engine  # noqa: F401
