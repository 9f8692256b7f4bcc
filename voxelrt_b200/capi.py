"""ctypes binding of include/voxelrt_b200.h (the C ABI of libvoxelrt_b200.so).

This is harness plumbing for tests/ and bench.py: the product is the shared library, whose
host side is C++ (voxelrt_b200/host/) exactly like the reference's.  There is no CPU
fallback — if the library is missing, `load()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = Path(__file__).resolve().parent / "lib" / "libvoxelrt_b200.so"

VRT_OK, VRT_ERR_INVALID, VRT_ERR_CUDA, VRT_ERR_OOM, VRT_ERR_STATE, VRT_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5  # VrtStatus
VRT_SECTOR_REMOVED = 1
VRT_HIT_NORMAL_MASK = 0x3F
VRT_HIT_HIT = 0x100
VRT_HIT_INBOUND = 0x200
VRT_HIT_CAPPED = 0x400
VRT_HIT_ITERS_SHIFT = 16
VRT_GLSL_COARSE = 1
VRT_GLSL_ANISOTROPIC = 2
VRT_FRAME_LINEAR_OUTPUT = 1
VRT_FRAME_AUX_HITS = 2
VRT_FRAME_COMPACT = 4
VRT_FRAME_PART_ROWS = 8
VRT_FRAME_GLSL = 16
VRT_FRAME_GLSL_ANISOTROPIC = 32
VRT_GATHER_DEPTH = 8
VRT_BLUE_NOISE_BYTES = 128 * 128 * 64 * 2


class VrtConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("sectors_xz_log2", C.c_uint32),
        ("sectors_y_log2", C.c_uint32),
        ("initial_brick_capacity", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class VrtDirtySector(C.Structure):
    _fields_ = [
        ("sx", C.c_int32),
        ("sy", C.c_int32),
        ("sz", C.c_int32),
        ("flags", C.c_uint32),
        ("alloc_mask", C.c_uint64),
        ("dirty_mask", C.c_uint64),
        ("bricks", C.c_void_p),
    ]


class VrtFrame(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("inv_proj", C.c_float * 16),
        ("proj", C.c_float * 16),
        ("world_origin", C.c_int32 * 3),
        ("origin_frac", C.c_float * 3),
        ("frame_no", C.c_uint32),
        ("bounces", C.c_uint32),
        ("max_iters", C.c_uint32),
        ("flags", C.c_uint32),
        ("part_index", C.c_uint32),
        ("part_count", C.c_uint32),
    ]


class VrtSkyDesc(C.Structure):
    _fields_ = [
        ("face_size", C.c_uint32),
        ("mip_levels", C.c_uint32),
        ("layer_shift", C.c_uint32),
        ("mip_offset", C.c_uint32 * 16),
        ("texel_count", C.c_uint64),
    ]


class VrtStats(C.Structure):
    _fields_ = [
        (n, C.c_uint64)
        for n in (
            "resident_bricks",
            "brick_capacity",
            "free_ranges",
            "resident_sectors",
            "bytes_uploaded",
            "bricks_uploaded",
            "bricks_relocated",
            "device_bytes",
            "last_launches",
            "bounce_form",
        )
    ]


class VrtTraversalMetrics(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "iters", "sector_fetches", "cell_fetches", "hits", "capped")]


# numpy views of the result records (must match the C structs byte for byte)
HIT_DTYPE = np.dtype(
    [
        ("vx", "<i4"),
        ("vy", "<i4"),
        ("vz", "<i4"),
        ("material", "<u4"),
        ("dist", "<f4"),
        ("px", "<f4"),
        ("py", "<f4"),
        ("pz", "<f4"),
        ("u", "<f4"),
        ("v", "<f4"),
        ("flags", "<u4"),
        ("_pad", "<u4"),
    ]
)
assert HIT_DTYPE.itemsize == 48
HITD_DTYPE = np.dtype(
    [
        ("dist", "<f8"),
        ("nx", "<f4"),
        ("ny", "<f4"),
        ("nz", "<f4"),
        ("u", "<f4"),
        ("v", "<f4"),
        ("vx", "<i4"),
        ("vy", "<i4"),
        ("vz", "<i4"),
        ("iters", "<u4"),
        ("_pad", "<u4"),
    ]
)
assert HITD_DTYPE.itemsize == 48
TILE_DTYPE = np.dtype([("albedo", "<u4", 16), ("depth", "<f4", 16), ("irr_rg", "<u4", 16), ("irr_bx", "<u4", 16)])
assert TILE_DTYPE.itemsize == 256
TILE_AD_DTYPE = np.dtype([("albedo", "<u4", 16), ("depth", "<f4", 16)])  # VrtTileAD (VRT_FRAME_COMPACT)
assert TILE_AD_DTYPE.itemsize == 128

# every symbol include/voxelrt_b200.h declares (tests check the .so exports them all)
EXPORTS = [
    "vrt_create",
    "vrt_destroy",
    "vrt_last_error",
    "vrt_get_stats",
    "vrt_set_palette",
    "vrt_sync",
    "vrt_read_sector",
    "vrt_trace",
    "vrt_trace_device",
    "vrt_trace_glsl",
    "vrt_hit_query",
    "vrt_set_blue_noise",
    "vrt_set_sky",
    "vrt_render",
    "vrt_render_device",
    "vrt_fb_export",
    "vrt_fb_import",
    "vrt_fb_release",
    "vrt_render_gather",
    "vrt_gather_wait",
    "vrt_get_metrics",
    "vrt_set_option",
]

_lib = None


def load(path: os.PathLike | None = None) -> C.CDLL:
    """Load libvoxelrt_b200.so (built by __graft_entry__.build()). Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise RuntimeError(
            f"{p} is missing: the CUDA library is not built (run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(str(p))
    vp, u32, u64, i32p = C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_int32)
    lib.vrt_create.argtypes = [C.POINTER(VrtConfig), C.POINTER(vp)]
    lib.vrt_destroy.argtypes = [vp]
    lib.vrt_destroy.restype = None
    lib.vrt_last_error.argtypes = [vp]
    lib.vrt_last_error.restype = C.c_char_p
    lib.vrt_get_stats.argtypes = [vp, C.POINTER(VrtStats)]
    lib.vrt_set_palette.argtypes = [vp, vp]
    lib.vrt_sync.argtypes = [vp, u32, C.POINTER(VrtDirtySector)]
    lib.vrt_read_sector.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(u64), C.POINTER(u32), vp, vp]
    lib.vrt_trace.argtypes = [vp, u64, vp, vp, i32p, u32, vp]
    lib.vrt_trace_device.argtypes = [vp, u64, vp, vp, i32p, u32, vp, vp]
    lib.vrt_trace_glsl.argtypes = [vp, u64, vp, vp, i32p, u32, vp]
    lib.vrt_hit_query.argtypes = [vp, u64, vp, vp, u32, vp]
    lib.vrt_set_blue_noise.argtypes = [vp, vp, C.c_size_t]
    lib.vrt_set_sky.argtypes = [vp, C.POINTER(VrtSkyDesc), vp]
    lib.vrt_render.argtypes = [vp, C.POINTER(VrtFrame), vp, vp]
    lib.vrt_render_device.argtypes = [vp, C.POINTER(VrtFrame), vp, vp, vp]
    lib.vrt_fb_export.argtypes = [vp, u64, vp, C.POINTER(vp)]
    lib.vrt_fb_import.argtypes = [vp, vp, C.POINTER(vp)]
    lib.vrt_fb_release.argtypes = [vp, vp]
    lib.vrt_render_gather.argtypes = [vp, C.POINTER(VrtFrame), vp, vp, vp]
    lib.vrt_gather_wait.argtypes = [vp, vp]
    lib.vrt_get_metrics.argtypes = [vp, C.POINTER(VrtTraversalMetrics)]
    lib.vrt_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    for name in EXPORTS:
        if name not in ("vrt_destroy", "vrt_last_error"):
            getattr(lib, name).restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def make_records(sectors):
    """sectors: iterable of (sx, sy, sz, alloc_mask, dirty_mask, bricks ndarray [k,512] u8 or None, removed)
    -> (ctypes array of VrtDirtySector, keep-alive list)."""
    sectors = list(sectors)
    arr = (VrtDirtySector * max(len(sectors), 1))()
    keep = []
    for i, s in enumerate(sectors):
        sx, sy, sz, alloc, dirty, bricks = s[:6]
        removed = s[6] if len(s) > 6 else False
        r = arr[i]
        r.sx, r.sy, r.sz = int(sx), int(sy), int(sz)
        r.flags = VRT_SECTOR_REMOVED if removed else 0
        r.alloc_mask = int(alloc)
        r.dirty_mask = int(dirty)
        if bricks is not None and len(bricks):
            b = np.ascontiguousarray(bricks, dtype=np.uint8)
            assert b.size == 512 * bin(int(alloc) & int(dirty)).count("1"), "brick payload must match popcount(dirty & alloc)"
            keep.append(b)
            r.bricks = b.ctypes.data
        else:
            r.bricks = None
    return arr, keep, len(sectors)


def make_frame(
    width,
    height,
    inv_proj,
    proj,
    world_origin,
    origin_frac,
    frame_no=1,
    bounces=0,
    max_iters=0,
    flags=0,
    part_index=0,
    part_count=1,
) -> VrtFrame:
    f = VrtFrame()
    f.width, f.height = int(width), int(height)
    ip = np.asarray(inv_proj, dtype=np.float32).reshape(16)
    pj = np.asarray(proj, dtype=np.float32).reshape(16)
    for i in range(16):
        f.inv_proj[i] = float(ip[i])
        f.proj[i] = float(pj[i])
    for i in range(3):
        f.world_origin[i] = int(world_origin[i])
        f.origin_frac[i] = float(np.float32(origin_frac[i]))
    f.frame_no, f.bounces, f.max_iters, f.flags = int(frame_no), int(bounces), int(max_iters), int(flags)
    f.part_index, f.part_count = int(part_index), int(part_count)
    return f


class VrtError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"vrt status {status}: {msg}")
        self.status = status


class Context:
    """Thin RAII wrapper over VrtContext*; mirrors the ABI one to one."""

    def __init__(self, sectors_xz_log2=6, sectors_y_log2=4, device=-1, initial_brick_capacity=0):
        self.lib = load()
        cfg = VrtConfig(C.sizeof(VrtConfig), device, sectors_xz_log2, sectors_y_log2, initial_brick_capacity, 0)
        self.h = C.c_void_p()
        st = self.lib.vrt_create(C.byref(cfg), C.byref(self.h))
        if st != 0:
            raise VrtError(st, (self.lib.vrt_last_error(None) or b"").decode())

    def _chk(self, st):
        if st != 0:
            raise VrtError(st, (self.lib.vrt_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.vrt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_palette(self, palette):
        p = np.ascontiguousarray(palette, dtype=np.uint64)
        assert p.size == 256
        self._chk(self.lib.vrt_set_palette(self.h, p.ctypes.data))

    def sync(self, sectors):
        arr, keep, n = make_records(sectors)
        self._chk(self.lib.vrt_sync(self.h, n, arr))
        del keep

    def sync_records(self, arr, n):
        """vrt_sync with a record array prepared by make_records() (keeps Python out of a timed loop)."""
        self._chk(self.lib.vrt_sync(self.h, n, arr))

    def read_sector(self, sx, sy, sz):
        mask, base = C.c_uint64(), C.c_uint32()
        bricks = np.zeros((64, 512), np.uint8)
        cells = np.zeros((64, 8), np.uint64)
        self._chk(self.lib.vrt_read_sector(self.h, sx, sy, sz, C.byref(mask), C.byref(base), bricks.ctypes.data, cells.ctypes.data))
        return mask.value, base.value, bricks, cells

    def stats(self) -> VrtStats:
        s = VrtStats()
        self._chk(self.lib.vrt_get_stats(self.h, C.byref(s)))
        return s

    def trace(self, origin3, dir3, world_origin, max_iters=0):
        o = np.ascontiguousarray(origin3, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dir3, dtype=np.float32).reshape(-1, 3)
        assert o.shape == d.shape
        out = np.zeros(o.shape[0], HIT_DTYPE)
        wo = (C.c_int32 * 3)(*[int(v) for v in world_origin])
        self._chk(self.lib.vrt_trace(self.h, o.shape[0], o.ctypes.data, d.ctypes.data, wo, max_iters, out.ctypes.data))
        return out

    def trace_glsl(self, origin3, dir3, world_origin, flags=0):
        """Ray casts with the GLSL renderer's semantics (rayCast / rayCastCoarse, VoxelTraversal.glsl:162-243)."""
        o = np.ascontiguousarray(origin3, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dir3, dtype=np.float32).reshape(-1, 3)
        assert o.shape == d.shape
        out = np.zeros(o.shape[0], HIT_DTYPE)
        wo = (C.c_int32 * 3)(*[int(v) for v in world_origin])
        self._chk(self.lib.vrt_trace_glsl(self.h, o.shape[0], o.ctypes.data, d.ctypes.data, wo, int(flags), out.ctypes.data))
        return out

    def trace_device(self, n, d_origin, d_dir, world_origin, max_iters, d_out, stream=0):
        wo = (C.c_int32 * 3)(*[int(v) for v in world_origin])
        self._chk(self.lib.vrt_trace_device(self.h, n, d_origin, d_dir, wo, max_iters, d_out, stream))

    def hit_query(self, origin3, dir3, max_iters=1024):
        o = np.ascontiguousarray(origin3, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(dir3, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(o.shape[0], HITD_DTYPE)
        self._chk(self.lib.vrt_hit_query(self.h, o.shape[0], o.ctypes.data, d.ctypes.data, max_iters, out.ctypes.data))
        return out

    def set_blue_noise(self, rg):
        b = np.ascontiguousarray(rg, dtype=np.uint8)
        self._chk(self.lib.vrt_set_blue_noise(self.h, b.ctypes.data, b.size))

    def set_sky(self, desc: VrtSkyDesc, texels):
        t = np.ascontiguousarray(texels, dtype=np.uint32)
        self._chk(self.lib.vrt_set_sky(self.h, C.byref(desc), t.ctypes.data))

    def render(self, frame: VrtFrame, want_aux=False):
        """Host-buffer render (the e2e path). Returns (out, aux) numpy arrays."""
        n = frame.width * frame.height
        compact = bool(frame.flags & VRT_FRAME_COMPACT)
        if frame.flags & VRT_FRAME_LINEAR_OUTPUT:
            out = np.zeros((2 if compact else 4, frame.height, frame.width), np.uint32)
        else:
            out = np.zeros(n // 16, TILE_AD_DTYPE if compact else TILE_DTYPE)
        aux = None
        if want_aux:
            frame.flags |= VRT_FRAME_AUX_HITS
            aux = np.zeros(n, HIT_DTYPE)
        self._chk(self.lib.vrt_render(self.h, C.byref(frame), out.ctypes.data, aux.ctypes.data if aux is not None else None))
        return out, aux

    def render_device(self, frame: VrtFrame, d_out, d_aux=None, stream=0):
        self._chk(self.lib.vrt_render_device(self.h, C.byref(frame), d_out, d_aux, stream))

    def render_gather(self, frame: VrtFrame, d_local, d_owner, stream=0):
        self._chk(self.lib.vrt_render_gather(self.h, C.byref(frame), d_local, d_owner, stream))

    def gather_wait(self, stream=0):
        self._chk(self.lib.vrt_gather_wait(self.h, stream))

    def fb_export(self, nbytes):
        handle = (C.c_uint8 * 64)()
        ptr = C.c_void_p()
        self._chk(self.lib.vrt_fb_export(self.h, nbytes, handle, C.byref(ptr)))
        return bytes(handle), ptr.value

    def fb_import(self, handle: bytes):
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        ptr = C.c_void_p()
        self._chk(self.lib.vrt_fb_import(self.h, buf, C.byref(ptr)))
        return ptr.value

    def fb_release(self, ptr):
        self._chk(self.lib.vrt_fb_release(self.h, ptr))

    def metrics(self) -> VrtTraversalMetrics:
        m = VrtTraversalMetrics()
        self._chk(self.lib.vrt_get_metrics(self.h, C.byref(m)))
        return m

    def set_option(self, name: str, value: int):
        self._chk(self.lib.vrt_set_option(self.h, name.encode(), int(value)))
