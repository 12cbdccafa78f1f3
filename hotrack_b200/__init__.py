"""hotrack_b200 -- sm_100a PointNet++ operator stack behind HOTrack's pointnet_lib API.

Importing the package loads libpn2b200.so (hand-written CUDA kernels, C ABI in
include/pn2b200.h).  There is no CPU fallback; a missing library is an ImportError.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is not built)

__version__ = "0.1.0"
