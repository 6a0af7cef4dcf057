"""amcl3d_test_b200 -- B200-native particle-weighting path of amcl3d behind the reference's API.

The product is libamcl3d_b200.so (hand-written sm_100a CUDA behind the C-ABI of
include/amcl3d_b200.h) plus the source-compatible C++ classes in include/amcl3d/.  This Python
package is the thin ctypes mirror of those classes that tests and bench.py drive; it holds no
compute of its own and has no CPU fallback.
"""
from .api import Grid3d, ParticleFilter, device_info  # noqa: F401
from .capi import Amcl3dError  # noqa: F401

__all__ = ["Grid3d", "ParticleFilter", "device_info", "Amcl3dError"]
