"""gaussiansplats3d_b200 -- B200 (sm_100a) depth -> sort -> rasterise engine behind the GaussianSplats3D
sort-worker / SplatMesh boundary.  Compute lives in csrc/libgsplat_b200.so (hand-written CUDA, C ABI in
include/gsplat_b200.h); this package is the host-side mirror of the reference's interface for that path."""
from . import _native
from ._native import GsError
from .engine import Engine, Uniforms, sort_indexes

__all__ = ["Engine", "Uniforms", "sort_indexes", "GsError", "_native"]
__version__ = "0.1.0"
