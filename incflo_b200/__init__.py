"""incflo_b200 -- B200-native approximate nodal projection (drop-in for incflo's
ApplyNodalProjection -> Hydro::NodalProjector -> MLMG/MLNodeLaplacian path).

csrc/   hand-written sm_100a CUDA kernels + the C ABI (include/b200np.h)
nodal_projector.py  host-side mirror of the reference interface (ctypes over the C ABI)
problems.py         synthetic inputs of the benchmark configurations
"""
from . import _lib  # noqa: F401
