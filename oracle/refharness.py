"""ctypes wrapper of oracle/_ref/libref_cpu.so — the reference's OWN CpuRenderer.cpp / VoxelMap.cpp
compiled from /root/reference by oracle/Makefile (see oracle/ref_harness.cpp).  TEST INFRASTRUCTURE:
pins the oracle restatement and serves as the `reference` CPU baseline in bench.py.  The library is
built with -march=native on an AVX-512 host; `available()` refuses to load it on a CPU without the
ISA it was built for."""
from __future__ import annotations

import ctypes as C
import time
from pathlib import Path

import numpy as np

from voxelrt_b200.capi import HIT_DTYPE, HITD_DTYPE, TILE_DTYPE, VrtDirtySector, VrtFrame, VrtSkyDesc, make_records

HERE = Path(__file__).resolve().parent
LIB = HERE / "_ref" / "libref_cpu.so"
# the same harness built with the flags the reference ships with (-O3 -march=native -ffast-math, src/CMakeLists.txt:36): used for
# TIMING only (bench.py's reference arm); every parity pin uses the -fno-fast-math build above
LIB_SHIPPED = HERE / "_ref" / "libref_cpu_fastmath.so"
# and with FP contraction off (-ffp-contract=off on the unmodified sources): the build the scene-ingest pin (ref_voxelize) compares with,
# because scenes/voxelizer.c states every fused multiply-add explicitly (none) instead of leaving the choice to the compiler
LIB_STRICT = HERE / "_ref" / "libref_cpu_strict.so"
_NEED = ("avx512f", "avx512bw", "avx512dq", "avx512vl", "avx2", "fma", "bmi2")
_libs = {}


def _cpu_ok() -> bool:
    try:
        flags = set()
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = set(line.split(":", 1)[1].split())
                break
        return all(f in flags for f in _NEED)
    except OSError:
        return False


def _path(shipped_flags):
    return LIB_STRICT if shipped_flags == "strict" else (LIB_SHIPPED if shipped_flags else LIB)


def available(shipped_flags=False) -> bool:
    return _path(shipped_flags).exists() and _cpu_ok()


def load(shipped_flags=False):
    if shipped_flags in _libs:
        return _libs[shipped_flags]
    if not available(shipped_flags):
        raise RuntimeError("oracle/_ref/libref_cpu*.so missing or this CPU lacks AVX-512")
    lib = C.CDLL(str(_path(shipped_flags)))
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    lib.ref_create.restype = vp
    lib.ref_destroy.argtypes = [vp]
    lib.ref_destroy.restype = None
    lib.ref_set_palette.argtypes = [vp, vp]
    lib.ref_set_palette.restype = None
    lib.ref_encode_material.argtypes = [C.c_uint8] * 4 + [C.c_float]
    lib.ref_encode_material.restype = u64
    lib.ref_sync.argtypes = [vp, u32, C.POINTER(VrtDirtySector)]
    lib.ref_read_sector.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(u64), vp, vp]
    lib.ref_set_blue_noise.argtypes = [vp]
    lib.ref_set_blue_noise.restype = None
    lib.ref_serialize.argtypes = [vp, C.c_char_p]
    lib.ref_deserialize.argtypes = [vp, C.c_char_p]
    lib.ref_set_material.argtypes = [vp, C.c_int] + [C.c_uint8] * 4 + [C.c_float]
    lib.ref_set_material.restype = None
    lib.ref_get_material.argtypes = [vp, C.c_int, vp, vp]
    lib.ref_get_material.restype = None
    lib.ref_map_sector_count.argtypes = [vp]
    lib.ref_map_sector_count.restype = u32
    lib.ref_map_list_sectors.argtypes = [vp, vp, vp, u32]
    lib.ref_map_list_sectors.restype = u32
    lib.ref_map_read_sector.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(u64), vp]
    lib.ref_set_sky.argtypes = [C.POINTER(VrtSkyDesc), vp]
    lib.ref_trace.argtypes = [vp, u64, vp, vp, C.POINTER(C.c_int32), vp, C.c_int]
    lib.ref_trace.restype = None
    lib.ref_hit_query.argtypes = [vp, u64, vp, vp, u32, vp]
    lib.ref_hit_query.restype = None
    lib.ref_primary_rays.argtypes = [C.POINTER(VrtFrame), vp, vp]
    lib.ref_primary_rays.restype = None
    lib.ref_render.argtypes = [vp, C.POINTER(VrtFrame), vp, C.c_int]
    lib.ref_render.restype = C.c_double
    lib.ref_inverse_proj_screen.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.ref_inverse_proj_screen.restype = None
    lib.ref_sincos_2pi.argtypes = [C.c_float, vp, vp]
    lib.ref_sincos_2pi.restype = None
    lib.ref_sample_direction.argtypes = [C.c_float, C.c_float, vp]
    lib.ref_sample_direction.restype = None
    lib.ref_blue_noise_tile.argtypes = [u32, u32, u32, u32, vp]
    lib.ref_blue_noise_tile.restype = None
    lib.ref_sky_sample.argtypes = [vp, C.c_int, vp]
    lib.ref_sky_sample.restype = None
    lib.ref_pack_r11g11b10f.argtypes = [C.c_float] * 3
    lib.ref_pack_r11g11b10f.restype = u32
    lib.ref_pack_rgba8.argtypes = [C.c_float] * 4
    lib.ref_pack_rgba8.restype = u32
    lib.ref_pack_rg16f.argtypes = [C.c_float] * 2
    lib.ref_pack_rg16f.restype = u32
    lib.ref_num_threads.restype = C.c_int
    lib.ref_voxelize.argtypes = [vp, u32, vp, vp, vp, u32, vp, vp, vp, u32]
    lib.ref_voxelize.restype = C.c_int
    lib.ref_brush_dispatch.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_float, C.c_int, C.c_uint8]
    lib.ref_brush_dispatch.restype = None
    lib.ref_map_take_dirty.argtypes = [vp, vp, vp, u32]
    lib.ref_map_take_dirty.restype = u32
    lib.ref_palette_build.argtypes = [vp, u64, u32, vp, vp]
    lib.ref_palette_build.restype = u32
    _libs[shipped_flags] = lib
    return lib


def palette_build(colors, max_colors=240, shipped_flags=False):
    """glim::PaletteBuilder: (palette uint8[n,3], FindIndex of every input colour)."""
    lib = load(shipped_flags)
    c = np.ascontiguousarray(colors, np.uint32).reshape(-1)
    rgb = np.zeros((256, 3), np.uint8)
    idx = np.zeros(max(1, c.size), np.uint8)
    n = lib.ref_palette_build(c.ctypes.data, c.size, int(max_colors), rgb.ctypes.data, idx.ctypes.data)
    return rgb[:n].copy(), idx[: c.size]


class RefMap:
    """The reference's VoxelMap + FlatVoxelStorage (a dense 2048x512x2048 view: ~2.3 GB of host memory)."""

    def __init__(self, shipped_flags=False):
        self.lib = load(shipped_flags)
        self.h = C.c_void_p(self.lib.ref_create())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.ref_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_palette(self, palette):
        p = np.ascontiguousarray(palette, dtype=np.uint64)
        self.lib.ref_set_palette(self.h, p.ctypes.data)

    def sync(self, sectors):
        arr, keep, n = make_records(sectors)
        self.lib.ref_sync(self.h, n, arr)
        del keep

    def read_sector(self, sx, sy, sz):
        mask = C.c_uint64()
        bricks = np.zeros((64, 512), np.uint8)
        cells = np.zeros((64, 8), np.uint64)
        st = self.lib.ref_read_sector(self.h, sx, sy, sz, C.byref(mask), bricks.ctypes.data, cells.ctypes.data)
        assert st == 0
        return mask.value, bricks, cells

    def set_blue_noise(self, rg):
        b = np.ascontiguousarray(rg, dtype=np.uint8)
        self.lib.ref_set_blue_noise(b.ctypes.data)

    # ---- the reference's own cvox (de)serialiser, VoxelMap.cpp:205-274 ----
    def serialize(self, path):
        if self.lib.ref_serialize(self.h, str(path).encode()) != 0:
            raise IOError("VoxelMap::Serialize failed")

    def deserialize(self, path):
        if self.lib.ref_deserialize(self.h, str(path).encode()) != 0:
            raise IOError("VoxelMap::Deserialize failed")

    def set_materials(self, mats):
        for i, (r, g, b, f, e) in enumerate(mats):
            self.lib.ref_set_material(self.h, i, int(r), int(g), int(b), int(f), float(e))

    def materials(self):
        out = []
        for i in range(256):
            rgbf = (C.c_uint8 * 4)()
            em = C.c_float()
            self.lib.ref_get_material(self.h, i, rgbf, C.byref(em))
            out.append((rgbf[0], rgbf[1], rgbf[2], rgbf[3], em.value))
        return out

    def brush(self, point_a, point_b, radius=30.0, action="fill", material=255):
        """BrushSession::Dispatch (Brush.cpp:10-37) on the VoxelMap; -> {(sx, sy, sz): dirty-brick mask} = VoxelMap::DirtyLocs, cleared."""
        a = (C.c_int32 * 3)(*[int(v) for v in point_a])
        b = (C.c_int32 * 3)(*[int(v) for v in point_b])
        self.take_dirty()
        self.lib.ref_brush_dispatch(self.h, a, b, float(radius), 1 if action == "replace" else 0, int(material))
        return self.take_dirty()

    def take_dirty(self):
        cap = 1 << 16
        xyz = np.zeros((cap, 3), np.int32)
        masks = np.zeros(cap, np.uint64)
        n = self.lib.ref_map_take_dirty(self.h, xyz.ctypes.data, masks.ctypes.data, cap)
        return {(int(xyz[i, 0]), int(xyz[i, 1]), int(xyz[i, 2])): int(masks[i]) for i in range(n)}

    def voxelize(self, tris, uvs, tri_tex, textures, size):
        """VoxelMap::VoxelizeModel(model, 0, size^3) on a decoded model (see ref_voxelize in ref_harness.cpp); voxels -> map_sectors(),
        palette -> materials()."""
        t = np.ascontiguousarray(tris, np.float32)
        u = np.ascontiguousarray(uvs, np.float32)
        tt = np.ascontiguousarray(tri_tex, np.int32)
        imgs = [np.ascontiguousarray(x, np.uint32) for x in textures]
        ptrs = (C.c_void_p * max(1, len(imgs)))(*[x.ctypes.data for x in imgs])
        tw = np.array([x.shape[1] for x in imgs] or [1], np.uint32)
        th = np.array([x.shape[0] for x in imgs] or [1], np.uint32)
        rc = self.lib.ref_voxelize(self.h, t.shape[0], t.ctypes.data, u.ctypes.data, tt.ctypes.data, len(imgs), ptrs, tw.ctypes.data, th.ctypes.data, int(size))
        if rc != 0:
            raise ValueError(f"ref_voxelize: {rc}")

    def map_sectors(self):
        """-> {(sx, sy, sz): (alloc_mask, bricks[k, 512])} read from the VoxelMap itself (not the renderer's dense view)."""
        n = self.lib.ref_map_sector_count(self.h)
        xyz = np.zeros((max(n, 1), 3), np.int32)
        masks = np.zeros(max(n, 1), np.uint64)
        n = self.lib.ref_map_list_sectors(self.h, xyz.ctypes.data, masks.ctypes.data, n)
        out = {}
        for i in range(n):
            k = bin(int(masks[i])).count("1")
            bricks = np.zeros((max(k, 1), 512), np.uint8)
            m = C.c_uint64()
            assert self.lib.ref_map_read_sector(self.h, int(xyz[i, 0]), int(xyz[i, 1]), int(xyz[i, 2]), C.byref(m), bricks.ctypes.data) == 0
            out[(int(xyz[i, 0]), int(xyz[i, 1]), int(xyz[i, 2]))] = (int(m.value), bricks[:k])
        return out

    def set_sky(self, desc, texels):
        t = np.ascontiguousarray(texels, dtype=np.uint32)
        if self.lib.ref_set_sky(C.byref(desc), t.ctypes.data) != 0:
            raise RuntimeError("sky layout disagrees with swr::Texture2D")

    def trace(self, origin3, dir3, world_origin, lanes_per_packet=16):
        """lanes_per_packet=1: lane-wise semantics (what the oracle restates); 16: packets as rendered."""
        o = np.ascontiguousarray(origin3, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dir3, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(o.shape[0], HIT_DTYPE)
        wo = (C.c_int32 * 3)(*[int(v) for v in world_origin])
        if o.shape[0]:
            self.lib.ref_trace(self.h, o.shape[0], o.ctypes.data, d.ctypes.data, wo, out.ctypes.data, lanes_per_packet)
        return out

    def hit_query(self, origin3, dir3, max_iters=1024):
        o = np.ascontiguousarray(origin3, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dir3, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(o.shape[0], HITD_DTYPE)
        self.lib.ref_hit_query(self.h, o.shape[0], o.ctypes.data, d.ctypes.data, max_iters, out.ctypes.data)
        return out

    def primary_rays(self, frame: VrtFrame):
        n = frame.width * frame.height
        o = np.zeros((n, 3), np.float32)
        d = np.zeros((n, 3), np.float32)
        self.lib.ref_primary_rays(C.byref(frame), o.ctypes.data, d.ctypes.data)
        return o, d

    def render(self, frame: VrtFrame, threads=0):
        """-> (tiles, seconds inside the row loop)."""
        nbytes = frame.width * frame.height * 16
        raw = np.zeros(nbytes + 64, np.uint8)  # Framebuffer::Tile is alignas(64) and stored with aligned moves
        off = (-raw.ctypes.data) % 64
        out = raw[off : off + nbytes].view(TILE_DTYPE)
        secs = self.lib.ref_render(self.h, C.byref(frame), out.ctypes.data, threads)
        return out, float(secs)


def host_threads() -> int:
    """Hardware threads this process may run on (not OMP_NUM_THREADS, which launchers override)."""
    import os

    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def encode_material(r, g, b, fuzz=255, emission=0.0):
    return int(load().ref_encode_material(r, g, b, fuzz, emission))


def inverse_proj_screen(m16, w, h):
    m = np.ascontiguousarray(m16, np.float32).reshape(16)
    out = np.zeros(16, np.float32)
    load().ref_inverse_proj_screen(m.ctypes.data, w, h, out.ctypes.data)
    return out


class RefRenderer:
    """bench.py helper: the reference renderer holding a scene; render_seconds() times one frame."""

    def __init__(self, scene, recs, shipped_flags: bool = False, threads: int = 0):
        self.map = RefMap(shipped_flags)
        self.map.set_palette(scene["palette"])
        self.map.sync(recs)
        self._sky = False
        self.shipped_flags = shipped_flags
        # explicit thread count: launchers such as torch.distributed.run export OMP_NUM_THREADS=1, which would silently
        # turn "all host threads" into one
        self.threads = int(threads) if threads else host_threads()

    def render_seconds(self, w, h, bounces):
        from scenes import camera, shading
        from voxelrt_b200 import capi

        if bounces and not self._sky:
            self.map.set_blue_noise(shading.load_blue_noise()[0])
            d, t, _ = shading.load_sky()
            self.map.set_sky(d, t)
            self._sky = True
        cam = camera.Camera()
        proj, inv, wo, frac = cam.matrices(w, h)
        frame = capi.make_frame(w, h, inv, proj, wo, frac, frame_no=1, bounces=bounces)
        _, secs = self.map.render(frame, threads=self.threads)
        return secs


def sincos_2pi(x):
    s, c = C.c_float(), C.c_float()
    load().ref_sincos_2pi(float(x), C.byref(s), C.byref(c))
    return np.float32(s.value), np.float32(c.value)


def sample_direction(sx, sy):
    out = np.zeros(3, np.float32)
    load().ref_sample_direction(float(sx), float(sy), out.ctypes.data)
    return out


def blue_noise_tile(x, y, frame_no, sample_idx):
    """-> float32[4,4,2]: rows = tile y, cols = tile x of the 4x4 tile containing (x, y)."""
    out = np.zeros(32, np.float32)
    load().ref_blue_noise_tile(x, y, frame_no, sample_idx, out.ctypes.data)
    return out.reshape(4, 4, 2)


def sky_sample(direction, mip):
    d = np.ascontiguousarray(direction, np.float32)
    out = np.zeros(3, np.float32)
    load().ref_sky_sample(d.ctypes.data, int(mip), out.ctypes.data)
    return out
