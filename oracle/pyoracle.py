"""ctypes wrapper of oracle/liboracle.so — the CPU ORACLE.  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
legs.  The product (voxelrt_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from voxelrt_b200.capi import (
    HIT_DTYPE,
    HITD_DTYPE,
    TILE_DTYPE,
    VRT_FRAME_LINEAR_OUTPUT,
    VrtDirtySector,
    VrtFrame,
    VrtSkyDesc,
    make_records,
)

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "iters", "sector_fetches", "cell_fetches", "hits", "capped")] + [
        ("iter_hist", C.c_uint64 * 8),
        ("lod_hist", C.c_uint64 * 6),
    ]

    def as_dict(self):
        d = {n: int(getattr(self, n)) for n in ("rays", "iters", "sector_fetches", "cell_fetches", "hits", "capped")}
        d["iter_hist"] = [int(v) for v in self.iter_hist]
        d["lod_hist"] = [int(v) for v in self.lod_hist]
        return d

    def algorithmic_bytes(self, primary_rays: int) -> int:
        """B = 8 I_s + 8 I_c + 9 H + 16 P."""
        return 8 * int(self.sector_fetches) + 8 * int(self.cell_fetches) + 9 * int(self.hits) + 16 * int(primary_rays)


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", str(HERE), "liboracle.so"], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB.exists():
        build()
    lib = C.CDLL(str(LIB))
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    lib.orc_map_create.argtypes = [u32, u32]
    lib.orc_map_create.restype = vp
    lib.orc_map_destroy.argtypes = [vp]
    lib.orc_map_destroy.restype = None
    lib.orc_map_set_palette.argtypes = [vp, vp]
    lib.orc_map_set_palette.restype = None
    lib.orc_map_sync.argtypes = [vp, u32, C.POINTER(VrtDirtySector)]
    lib.orc_build_occupancy.argtypes = [vp, vp]
    lib.orc_build_occupancy.restype = None
    lib.orc_map_read_sector.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(u64), vp, vp]
    lib.orc_set_blue_noise.argtypes = [vp, vp, C.c_size_t]
    lib.orc_set_blue_noise.restype = None
    lib.orc_set_sky.argtypes = [vp, C.POINTER(VrtSkyDesc), vp]
    lib.orc_set_sky.restype = None
    lib.orc_trace.argtypes = [vp, u64, vp, vp, C.POINTER(C.c_int32), u32, vp, C.POINTER(OrcStats), C.c_int]
    lib.orc_trace.restype = None
    lib.orc_trace_glsl.argtypes = [vp, u64, vp, vp, C.POINTER(C.c_int32), u32, vp, C.POINTER(OrcStats), C.c_int]
    lib.orc_trace_glsl.restype = None
    lib.orc_render_glsl.argtypes = [vp, C.POINTER(VrtFrame), vp, C.POINTER(OrcStats), C.c_int]
    lib.orc_render_glsl.restype = None
    lib.orc_interaction_lut.argtypes = [vp]
    lib.orc_interaction_lut.restype = None
    lib.orc_hit_query.argtypes = [vp, u64, vp, vp, u32, vp, C.c_int]
    lib.orc_hit_query.restype = None
    lib.orc_render.argtypes = [vp, C.POINTER(VrtFrame), vp, vp, C.POINTER(OrcStats), C.c_int, u32, u32]
    lib.orc_render.restype = None
    lib.orc_primary_ray.argtypes = [C.POINTER(VrtFrame), u32, u32, vp, vp]
    lib.orc_primary_ray.restype = None
    lib.orc_sample_direction.argtypes = [C.c_float, C.c_float, vp]
    lib.orc_sample_direction.restype = None
    lib.orc_blue_noise_sample.argtypes = [vp, u32, u32, u32, u32, vp]
    lib.orc_blue_noise_sample.restype = None
    lib.orc_sky_sample.argtypes = [vp, vp, u32, vp]
    lib.orc_sky_sample.restype = None
    lib.orc_sincos_2pi.argtypes = [C.c_float, vp, vp]
    lib.orc_sincos_2pi.restype = None
    lib.orc_pack_r11g11b10f.argtypes = [C.c_float, C.c_float, C.c_float]
    lib.orc_pack_r11g11b10f.restype = u32
    lib.orc_encode_material.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint8, C.c_float]
    lib.orc_encode_material.restype = u64
    lib.orc_num_threads.restype = C.c_int
    lib.orc_debug_pixel.argtypes = [vp, C.POINTER(VrtFrame), u32, u32, vp, vp, vp]
    lib.orc_debug_pixel.restype = u32
    _lib = lib
    return lib


class OracleMap:
    def __init__(self, sectors_xz_log2=6, sectors_y_log2=4):
        self.lib = load()
        self.h = C.c_void_p(self.lib.orc_map_create(sectors_xz_log2, sectors_y_log2))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.orc_map_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_palette(self, palette):
        p = np.ascontiguousarray(palette, dtype=np.uint64)
        self.lib.orc_map_set_palette(self.h, p.ctypes.data)

    def sync(self, sectors):
        arr, keep, n = make_records(sectors)
        self.lib.orc_map_sync(self.h, n, arr)
        del keep

    def read_sector(self, sx, sy, sz):
        mask = C.c_uint64()
        bricks = np.zeros((64, 512), np.uint8)
        cells = np.zeros((64, 8), np.uint64)
        self.lib.orc_map_read_sector(self.h, sx, sy, sz, C.byref(mask), bricks.ctypes.data, cells.ctypes.data)
        return mask.value, bricks, cells

    def set_blue_noise(self, rg):
        b = np.ascontiguousarray(rg, dtype=np.uint8)
        self.lib.orc_set_blue_noise(self.h, b.ctypes.data, b.size)

    def set_sky(self, desc, texels):
        t = np.ascontiguousarray(texels, dtype=np.uint32)
        self.lib.orc_set_sky(self.h, C.byref(desc), t.ctypes.data)

    def trace(self, origin3, dir3, world_origin, max_iters=0, threads=0):
        o = np.ascontiguousarray(origin3, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dir3, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(o.shape[0], HIT_DTYPE)
        st = OrcStats()
        wo = (C.c_int32 * 3)(*[int(v) for v in world_origin])
        self.lib.orc_trace(self.h, o.shape[0], o.ctypes.data, d.ctypes.data, wo, max_iters, out.ctypes.data, C.byref(st), threads)
        return out, st

    def trace_glsl(self, origin3, dir3, world_origin, flags=0, threads=0):
        """rayCast / rayCastCoarse of the GLSL renderer (VoxelTraversal.glsl:162-243); flags = VRT_GLSL_*."""
        o = np.ascontiguousarray(origin3, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dir3, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(o.shape[0], HIT_DTYPE)
        st = OrcStats()
        wo = (C.c_int32 * 3)(*[int(v) for v in world_origin])
        self.lib.orc_trace_glsl(self.h, o.shape[0], o.ctypes.data, d.ctypes.data, wo, int(flags), out.ctypes.data, C.byref(st), threads)
        return out, st

    def hit_query(self, origin3, dir3, max_iters=1024, threads=0):
        o = np.ascontiguousarray(origin3, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dir3, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(o.shape[0], HITD_DTYPE)
        self.lib.orc_hit_query(self.h, o.shape[0], o.ctypes.data, d.ctypes.data, max_iters, out.ctypes.data, threads)
        return out

    def render(self, frame: VrtFrame, want_aux=False, threads=0, row0=0, row1=None):
        n = frame.width * frame.height
        if frame.flags & VRT_FRAME_LINEAR_OUTPUT:
            out = np.zeros((4, frame.height, frame.width), np.uint32)
        else:
            out = np.zeros(n // 16, TILE_DTYPE)
        aux = np.zeros(n, HIT_DTYPE) if want_aux else None
        st = OrcStats()
        self.lib.orc_render(
            self.h,
            C.byref(frame),
            out.ctypes.data,
            aux.ctypes.data if aux is not None else None,
            C.byref(st),
            threads,
            row0,
            frame.height if row1 is None else row1,
        )
        return out, aux, st

    def sky_sample(self, dir3, mip):
        """CpuRenderer's SkyBox.SampleCube at an integer mip, times 3 (CpuRenderer.cpp:348-362): float32[3]."""
        d = np.ascontiguousarray(dir3, np.float32)
        out = np.zeros(3, np.float32)
        self.lib.orc_sky_sample(self.h, d.ctypes.data, int(mip), out.ctypes.data)
        return out

    def sincos_2pi(self, x):
        """simd::sincos_2pi (SIMD.h:175-190): (sin, cos) of 2 pi x as float32."""
        s, c = C.c_float(), C.c_float()
        self.lib.orc_sincos_2pi(float(x), C.byref(s), C.byref(c))
        return np.float32(s.value), np.float32(c.value)

    def render_glsl(self, frame: VrtFrame, threads=0):
        """A VRT_FRAME_GLSL frame (the GPU renderer's VoxelRender.comp per pixel) -> (tiles or 4 planes, stats)."""
        n = frame.width * frame.height
        out = np.zeros((4, frame.height, frame.width), np.uint32) if frame.flags & VRT_FRAME_LINEAR_OUTPUT else np.zeros(n // 16, TILE_DTYPE)
        st = OrcStats()
        self.lib.orc_render_glsl(self.h, C.byref(frame), out.ctypes.data, C.byref(st), threads)
        return out, st

    def debug_pixel(self, frame: VrtFrame, x, y):
        """-> (rays[n,6], hits[n], out4) of everything pixel (x,y) casts."""
        rays = np.zeros((8, 6), np.float32)
        hits = np.zeros(8, HIT_DTYPE)
        out4 = np.zeros(4, np.uint32)
        n = self.lib.orc_debug_pixel(self.h, C.byref(frame), x, y, rays.ctypes.data, hits.ctypes.data, out4.ctypes.data)
        return rays[:n], hits[:n], out4

    def primary_rays(self, frame: VrtFrame):
        """(origins, dirs) of every pixel, row-major — GetPrimaryRay through the oracle."""
        w, h = frame.width, frame.height
        o = np.zeros((h, w, 3), np.float32)
        d = np.zeros((h, w, 3), np.float32)
        oo = (C.c_float * 3)()
        dd = (C.c_float * 3)()
        for y in range(h):
            for x in range(w):
                self.lib.orc_primary_ray(C.byref(frame), x, y, oo, dd)
                o[y, x] = oo[:]
                d[y, x] = dd[:]
        return o.reshape(-1, 3), d.reshape(-1, 3)


def build_occupancy(brick):
    lib = load()
    b = np.ascontiguousarray(brick, dtype=np.uint8).reshape(512)
    out = np.zeros(8, np.uint64)
    lib.orc_build_occupancy(b.ctypes.data, out.ctypes.data)
    return out


def encode_material(r, g, b, fuzz=255, emission=0.0):
    return int(load().orc_encode_material(r, g, b, fuzz, emission))


def pack_r11g11b10f(r, g, b):
    return int(load().orc_pack_r11g11b10f(r, g, b))


def num_threads():
    return int(load().orc_num_threads())


def interaction_lut() -> np.ndarray:
    """GenerateRayCellInteractionMaskLUT (GpuRenderer.cpp:193-210) as the oracle restates it: 512 u64."""
    t = np.zeros(512, np.uint64)
    load().orc_interaction_lut(t.ctypes.data)
    return t
