"""voxelrt_b200 — B200-native (sm_100a) brickmap traversal path of dubiousconst282/VoxelRT.

The product is `lib/libvoxelrt_b200.so` (hand-written CUDA + C++ host, C ABI in
include/voxelrt_b200.h).  This package holds its sources (`csrc/`, `host/`) and the ctypes
binding the tests and the bench harness use (`capi`).  No CPU fallback exists.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
