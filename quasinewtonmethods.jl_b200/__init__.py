"""quasinewtonmethods.jl_b200 — B200-native (sm_100a) BFGS engine.

A drop-in for the hot path of chriselrod/QuasiNewtonMethods.jl
(`BFGSState` / `optimize!` / `optimum` / `gradient` / `BackTracking` and the
`logdensity` / `∂logdensity!` objective protocol) plus a batched `optimize!`.
All numerics live in libqn_cuda.so (hand-written CUDA, C ABI in
include/qn_cuda.h); this package is the Python host mirror of the reference's
interface.  The directory name contains a dot, so import it through the alias
module at the repo root:  `import qn_b200`.
"""
from . import _lib, sharding
from .api import *  # noqa: F401,F403
from .api import __all__  # noqa: F401

__version__ = "0.1.0"
