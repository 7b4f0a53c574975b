"""boss_b200 -- B200-native (sm_100a) backend for the GP hot path of soldasim/BOSS.jl.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (include/boss_b200.h) -> lib/libboss_b200.so
  _lib.py    ctypes binding (no CPU fallback)
  the remaining modules mirror the reference's Julia interfaces for this path (same names, argument
  meaning and error behaviour) on top of the C ABI.
"""
from . import _lib  # noqa: F401  (raises loudly when the CUDA library has not been built)

__all__ = ["_lib"]
