"""minsu3d_b200 -- B200-native (sm_100a) sparse-voxel hot path behind PointGroup / HAIS / SoftGroup.

Layout (only what the path needs):
    csrc/            hand-written CUDA kernels + the C ABI (include/b2s.h) -> lib/libb2s.so
    _cabi.py, ops.py ctypes binding and tensor-level wrappers (torch = memory + streams)
    MinkowskiEngine/ drop-in for the ME Python surface minsu3d uses          (boundary #1)
    COMMON_OPS.py    drop-in for the reference's pybind module               (boundary #2)
    common_ops/      mirror of minsu3d/common_ops/functions/*.py
    harness/         synthetic scenes + the reference's model code paths that call the hot path
    dp.py            scene-sharded data parallelism (bucketed NCCL gradient all-reduce)
"""
import sys

__version__ = "0.1.0"


def install_as_reference_modules():
    """Make `import MinkowskiEngine` / `import COMMON_OPS` resolve to this package, so the
    reference's own Python (minsu3d/model/**, minsu3d/common_ops/functions/*.py) runs unmodified."""
    from . import COMMON_OPS, MinkowskiEngine
    sys.modules["MinkowskiEngine"] = MinkowskiEngine
    sys.modules["MinkowskiEngine.utils"] = MinkowskiEngine.utils
    sys.modules["COMMON_OPS"] = COMMON_OPS
    return MinkowskiEngine, COMMON_OPS
