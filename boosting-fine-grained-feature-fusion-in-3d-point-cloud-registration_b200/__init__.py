"""kpreg_b200 — the KPConv registration hot path of RegTR / "Boosting Fine-grained Feature Fusion in
3D Point Cloud Registration" on B200 (sm_100a): voxel-grid subsampling pyramid, radius-neighbour
tables, KPConv and weighted Kabsch as hand-written CUDA kernels behind a C ABI
(``include/kpreg_b200.h``), exposed through the reference's own Python API.

    from kpreg_b200.cpp_wrappers import cpp_subsampling, cpp_neighbors
    from kpreg_b200.kpconv import Preprocessor, KPFEncoder
    from kpreg_b200.kpconv_blocks import KPConv
    from kpreg_b200.se3_torch import compute_rigid_transform, fast_compute_rigid_transform
"""
from . import _lib  # noqa: F401  (ctypes binding; the library is loaded on first use)
from .config import AttrDict, kpconv_config  # noqa: F401



def invalidate_caches() -> None:
    """Drop the inference caches of derived weights (see ops.invalidate_caches): call after writing parameters or
    BatchNorm statistics through ``.data`` — such writes do not bump the version counters the caches are checked by."""
    from . import ops
    ops.invalidate_caches()


__all__ = ["AttrDict", "kpconv_config", "invalidate_caches"]
__version__ = "0.1.0"
