"""ctypes wrapper of the image-space part of oracle/liboracle.so (oracle/vrt_post_oracle.c) — the CPU ORACLE of
the reference's GBuffer shaders.  TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's CPU legs
may import it; the product (voxelrt_b200/) never does.  Parity unpinned (no GL device here; see the C file's header).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import pyoracle

_bound = False


def load():
    global _bound
    lib = pyoracle.load()
    if not _bound:
        vp = C.c_void_p
        lib.post_oracle_create.argtypes = [C.c_int, C.c_int]
        lib.post_oracle_create.restype = vp
        lib.post_oracle_destroy.argtypes = [vp]
        lib.post_oracle_destroy.restype = None
        lib.post_oracle_set_camera.argtypes = [vp, vp, vp, vp]
        lib.post_oracle_set_camera.restype = None
        lib.post_oracle_frame.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
        lib.post_oracle_frame.restype = None
        lib.post_oracle_read.argtypes = [vp, C.c_int, vp]
        lib.post_oracle_read.restype = C.c_int
        lib.post_exp.argtypes = [C.c_float]
        lib.post_exp.restype = C.c_float
        lib.post_log.argtypes = [C.c_float]
        lib.post_log.restype = C.c_float
        lib.post_f32_to_f16.argtypes = [C.c_float]
        lib.post_f32_to_f16.restype = C.c_uint16
        lib.post_f16_to_f32.argtypes = [C.c_uint16]
        lib.post_f16_to_f32.restype = C.c_float
        _bound = True
    return lib


class PostOracle:
    """Texture-per-texture restatement of GBuffer (GBuffer.h) + its shaders; same call order as the product's GBuffer."""

    IRR, PREV_IRR, TEMP_IRR, MOMENTS, HIST, DEPTH, ALBEDO = range(7)

    def __init__(self, width: int, height: int):
        self.lib = load()
        self.w, self.h = int(width), int(height)
        self.o = C.c_void_p(self.lib.post_oracle_create(self.w, self.h))
        self.passes, self.channel, self._reset = 5, 0, 0

    def __del__(self):
        try:
            self.lib.post_oracle_destroy(self.o)
        except Exception:
            pass

    def set_passes(self, n):
        self.passes = int(n)

    def set_debug_channel(self, ch):
        self.channel = int(ch)

    def set_camera(self, proj, inv_proj, position, reset_history=False):
        pj = np.ascontiguousarray(np.asarray(proj, np.float32).reshape(16))
        ip = np.ascontiguousarray(np.asarray(inv_proj, np.float32).reshape(16))
        ps = np.ascontiguousarray(np.asarray(position, np.float64).reshape(3))
        self.lib.post_oracle_set_camera(self.o, pj.ctypes.data, ip.ctypes.data, ps.ctypes.data)
        self._reset = 1 if reset_history else 0

    def denoise_present(self, tiles: np.ndarray) -> np.ndarray:
        t = np.ascontiguousarray(tiles).view(np.uint8).reshape(-1)
        assert t.size == self.w * self.h * 16
        out = np.empty((self.h, self.w), np.uint32)
        self.lib.post_oracle_frame(self.o, t.ctypes.data, self._reset, self.passes, self.channel, out.ctypes.data)
        return out

    def read(self, which: int) -> np.ndarray:
        n = self.w * self.h
        shape_dtype = {
            self.IRR: ((n, 4), np.uint16),
            self.PREV_IRR: ((n, 4), np.uint16),
            self.TEMP_IRR: ((n, 4), np.uint16),
            self.MOMENTS: ((n, 2), np.uint16),
            self.HIST: ((n,), np.uint8),
            self.DEPTH: ((n,), np.float32),
            self.ALBEDO: ((n,), np.uint32),
        }[which]
        out = np.empty(*shape_dtype)
        assert self.lib.post_oracle_read(self.o, int(which), out.ctypes.data) == 0
        return out
