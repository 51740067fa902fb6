"""B200-native hierarchical block matching with 8-connected regularisation.

Drop-in for one path of ashish-nr/BlockBasedMotionEstimation: `MF` (motion_framework) driving
`PyramidLevel` / `BlockPosition`, and `.flo` I/O through `Flow` (rw_flow).  All compute runs in hand-written
sm_100a CUDA kernels behind the C ABI of libbbme.so (include/bbme.h); this package is the Python mirror of
the reference's C++ interface used by the tests and the benchmark.
"""
from .api import MF, Flow, PyramidLevel, BlockPosition, Estimator, Pool, BbmeError, plan_shape  # noqa: F401

__all__ = ["MF", "Flow", "PyramidLevel", "BlockPosition", "Estimator", "Pool", "BbmeError", "plan_shape"]
