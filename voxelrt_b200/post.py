"""ctypes binding of include/voxelrt_b200_post.h — the reference's `GBuffer` (src/VoxelRT/GBuffer.h) on the GPU:
blit of the tiled framebuffer, temporal reprojection, SVGF variance + à-trous passes, tone-mapped present.

Harness plumbing for tests/ and bench.py like capi.py; the product is the shared library.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

VRT_CHANNEL_NONE, VRT_CHANNEL_ALBEDO, VRT_CHANNEL_IRRADIANCE, VRT_CHANNEL_NORMALS = 0, 1, 2, 3
VRT_CHANNEL_TRAVERSAL_ITERS, VRT_CHANNEL_VARIANCE = 4, 5
VRT_PLANE_IRRADIANCE, VRT_PLANE_PREV_IRRADIANCE, VRT_PLANE_TEMP_IRRADIANCE, VRT_PLANE_MOMENTS, VRT_PLANE_HISTORY_LEN = range(5)

# one filter record: the IrradianceTex texel with the DepthTex / AlbedoTex texels of the same frame next to it
RECORD_DTYPE = np.dtype([("irr", "<u2", 4), ("depth", "<f4"), ("albedo", "<u4")])
assert RECORD_DTYPE.itemsize == 16


class VrtGBufferCamera(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("proj", C.c_float * 16),
        ("inv_proj", C.c_float * 16),
        ("position", C.c_double * 3),
        ("reset_history", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


# every symbol include/voxelrt_b200_post.h declares
EXPORTS = [
    "vrt_gbuffer_create",
    "vrt_gbuffer_destroy",
    "vrt_gbuffer_last_error",
    "vrt_gbuffer_set_passes",
    "vrt_gbuffer_set_debug_channel",
    "vrt_gbuffer_set_camera",
    "vrt_gbuffer_denoise_present",
    "vrt_gbuffer_denoise_present_device",
    "vrt_gbuffer_render_present",
    "vrt_gbuffer_read",
    "vrt_gbuffer_last_launches",
]

_bound = False


def load() -> C.CDLL:
    global _bound
    lib = capi.load()
    if not _bound:
        vp, u32 = C.c_void_p, C.c_uint32
        lib.vrt_gbuffer_create.argtypes = [C.c_int32, C.POINTER(vp)]
        lib.vrt_gbuffer_destroy.argtypes = [vp]
        lib.vrt_gbuffer_destroy.restype = None
        lib.vrt_gbuffer_last_error.argtypes = [vp]
        lib.vrt_gbuffer_last_error.restype = C.c_char_p
        lib.vrt_gbuffer_set_passes.argtypes = [vp, u32]
        lib.vrt_gbuffer_set_debug_channel.argtypes = [vp, u32]
        lib.vrt_gbuffer_set_camera.argtypes = [vp, C.POINTER(VrtGBufferCamera)]
        lib.vrt_gbuffer_denoise_present.argtypes = [vp, vp, vp]
        lib.vrt_gbuffer_denoise_present_device.argtypes = [vp, vp, vp, vp]
        lib.vrt_gbuffer_render_present.argtypes = [vp, vp, C.POINTER(capi.VrtFrame), vp]
        lib.vrt_gbuffer_read.argtypes = [vp, u32, vp]
        lib.vrt_gbuffer_last_launches.argtypes = [vp, C.POINTER(C.c_uint64)]
        for name in EXPORTS:
            if name not in ("vrt_gbuffer_destroy", "vrt_gbuffer_last_error"):
                getattr(lib, name).restype = C.c_int
        _bound = True
    return lib


def make_camera(width, height, proj, inv_proj, position, reset_history=False) -> VrtGBufferCamera:
    c = VrtGBufferCamera()
    c.width, c.height = int(width), int(height)
    pj = np.asarray(proj, dtype=np.float32).reshape(16)
    ip = np.asarray(inv_proj, dtype=np.float32).reshape(16)
    for i in range(16):
        c.proj[i] = float(pj[i])
        c.inv_proj[i] = float(ip[i])
    for i in range(3):
        c.position[i] = float(position[i])
    c.reset_history = 1 if reset_history else 0
    return c


class GBuffer:
    """The reference's GBuffer object: SetCamera per frame, then DenoiseAndPresent on the traced tiles."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.vrt_gbuffer_create(int(device), C.byref(h))
        if st != 0:
            raise capi.VrtError(st, (self.lib.vrt_gbuffer_last_error(None) or b"").decode())
        self.h = h
        self.width = self.height = 0

    def _chk(self, st):
        if st != 0:
            raise capi.VrtError(st, (self.lib.vrt_gbuffer_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.vrt_gbuffer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_passes(self, n):
        self._chk(self.lib.vrt_gbuffer_set_passes(self.h, int(n)))

    def set_debug_channel(self, ch):
        self._chk(self.lib.vrt_gbuffer_set_debug_channel(self.h, int(ch)))

    def set_camera(self, cam: VrtGBufferCamera):
        self._chk(self.lib.vrt_gbuffer_set_camera(self.h, C.byref(cam)))
        self.width, self.height = cam.width, cam.height

    def denoise_present(self, tiles: np.ndarray) -> np.ndarray:
        """tiles: the w*h*16-byte tile framebuffer (host).  Returns the presented image, (h, w) u32 RGBA8."""
        t = np.ascontiguousarray(tiles).view(np.uint8).reshape(-1)
        assert t.size == self.width * self.height * 16, "tile framebuffer size does not match the camera's view size"
        out = np.empty((self.height, self.width), np.uint32)
        self._chk(self.lib.vrt_gbuffer_denoise_present(self.h, t.ctypes.data, out.ctypes.data))
        return out

    def denoise_present_device(self, d_tiles: int, d_out: int, stream: int = 0):
        self._chk(self.lib.vrt_gbuffer_denoise_present_device(self.h, C.c_void_p(d_tiles), C.c_void_p(d_out), C.c_void_p(stream)))

    def render_present(self, ctx: "capi.Context", frame: "capi.VrtFrame") -> np.ndarray:
        """RenderFrame from the trace to the window: trace on `ctx`, denoise, present; only the RGBA8 image comes back."""
        out = np.empty((self.height, self.width), np.uint32)
        self._chk(self.lib.vrt_gbuffer_render_present(self.h, ctx.h, C.byref(frame), out.ctypes.data))
        return out

    def read(self, plane: int) -> np.ndarray:
        n = self.width * self.height
        if plane in (VRT_PLANE_IRRADIANCE, VRT_PLANE_PREV_IRRADIANCE, VRT_PLANE_TEMP_IRRADIANCE):
            out = np.empty(n, RECORD_DTYPE)
        elif plane == VRT_PLANE_MOMENTS:
            out = np.empty((n, 2), np.uint16)
        else:
            out = np.empty(n, np.uint8)
        self._chk(self.lib.vrt_gbuffer_read(self.h, int(plane), out.ctypes.data))
        return out

    def last_launches(self) -> int:
        v = C.c_uint64()
        self._chk(self.lib.vrt_gbuffer_last_launches(self.h, C.byref(v)))
        return int(v.value)
