"""B200-native LightGlue matcher (drop-in for gluefactory.models.matchers.lightglue).

Only the hot path lives here: `lightglue.LightGlue` (host mirror of the reference
plugin), `_abi` (ctypes binding of include/lightglue_b200.h), `csrc/` (CUDA
kernels), `build` (nvcc driver) and `synthetic` (seeded inputs for tests/bench).
"""
from .lightglue import LightGlue  # noqa: F401

__all__ = ["LightGlue"]
