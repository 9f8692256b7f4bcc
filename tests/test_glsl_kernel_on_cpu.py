"""voxelrt_b200/csrc/vrt_glsl.cuh (the kernels behind vrt_trace_glsl), compiled for the host behind tests/native/cuda_host_shim.h and run
thread by thread over device-layout arrays built here with numpy (bordered header grid, brick slots, cell masks, voxels), against the
oracle's orc_trace_glsl — every VrtHit field bit-exact, in the CPU tier.  Also exercises the layout contract of DESIGN.md §4 from the
outside: headers {allocLo, allocHi, baseSlot, baseSlot + popc(allocLo)}, slot = base + rank of the brick bit."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import assert_hits_equal, random_rays
from test_glsl_oracle import camera_frame_rays

NATIVE = Path(__file__).resolve().parent / "native"


class EmuScene(C.Structure):
    _fields_ = [("hdr", C.c_void_p), ("cells", C.c_void_p), ("voxels", C.c_void_p), ("palette", C.c_void_p), ("sxz", C.c_uint32), ("sy", C.c_uint32)]


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", str(NATIVE), "libemu_glsl.so"], check=True)
    lib = C.CDLL(str(NATIVE / "libemu_glsl.so"))
    lib.emu_build_groups.argtypes = [C.POINTER(EmuScene), C.c_void_p]
    lib.emu_build_groups.restype = None
    lib.emu_trace_glsl.argtypes = [C.POINTER(EmuScene), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p]
    lib.emu_trace_glsl.restype = None
    return lib


def aligned_zeros(shape, dtype, align=256):
    """cudaMalloc returns 256-byte aligned memory and the fast traversal loop relies on it (it ORs the 8 * cell offset into the
    address of a brick's 64-byte cell-mask block); numpy only promises 16 bytes."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off : off + n].view(dtype).reshape(shape)


class DeviceLayout:
    """The resident brickmap as the kernels see it, built from a scene dict without any product code."""

    def __init__(self, scene, sxz=6, sy=4):
        from glsl_cast_model_py import interaction_mask

        self.sxz, self.sy = sxz, sy
        sxp, syp = (1 << sxz) + 2, (1 << sy) + 2
        # k_init_headers: the bordered grid with 2 * sxp^2 OUTSIDE guard entries before and after it (the fast traversal loop does
        # not clamp its header index)
        n, guard = sxp * sxp * syp, 2 * sxp * sxp
        self.hdr_all = aligned_zeros((n + 2 * guard, 4), np.uint32)
        self.hdr_all[:, 3] = 0x80000000  # VRT_HDR_OUTSIDE
        self.hdr = self.hdr_all[guard : guard + n]  # a view: entry 0 of the grid
        inside = np.zeros((syp, sxp, sxp), bool)  # [y, z, x] of the bordered grid
        inside[1:-1, 1:-1, 1:-1] = True
        self.hdr[inside.reshape(-1), 3] = 0
        n_bricks = sum(bin(int(m)).count("1") for m, _ in scene["sectors"].values())
        self.cells = aligned_zeros((n_bricks * 8, 2), np.uint32)
        self.voxels = aligned_zeros(n_bricks * 512, np.uint8)
        slot = 0
        weights = (np.uint64(1) << (np.arange(4, dtype=np.uint64)[None, None, :] + np.uint64(4) * np.arange(4, dtype=np.uint64)[None, :, None]
                                    + np.uint64(16) * np.arange(4, dtype=np.uint64)[:, None, None]))  # [y, z, x] -> bit x + 4 z + 16 y
        for (sx, sy_, sz), (mask, bricks) in sorted(scene["sectors"].items()):
            mask = int(mask)
            lo, hi = mask & 0xFFFFFFFF, mask >> 32
            self.hdr[(sx + 1) + (sz + 1) * sxp + (sy_ + 1) * sxp * sxp] = (lo, hi, slot, slot + bin(lo).count("1"))
            k = bin(mask).count("1")
            vox = np.asarray(bricks, np.uint8).reshape(k, 8, 8, 8)  # [brick, y, z, x]
            self.voxels[slot * 512 : (slot + k) * 512] = vox.reshape(-1)
            # cell c = cx | cz<<1 | cy<<2 holds voxels [4cy:4cy+4, 4cz:4cz+4, 4cx:4cx+4]
            sub = (vox != 0).reshape(k, 2, 4, 2, 4, 2, 4).transpose(0, 1, 3, 5, 2, 4, 6)  # [brick, cy, cz, cx, y, z, x]
            bits = (sub.astype(np.uint64) * weights).sum(axis=(4, 5, 6), dtype=np.uint64).reshape(k, 8)  # index cy*4 + cz*2 + cx
            cells = self.cells[slot * 8 : (slot + k) * 8].reshape(k, 8, 2)
            cells[..., 0] = (bits & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            cells[..., 1] = (bits >> np.uint64(32)).astype(np.uint32)
            slot += k
        pal = np.asarray(scene["palette"], np.uint64)
        self.palette = np.stack([(pal & np.uint64(0xFFFFFFFF)).astype(np.uint32), (pal >> np.uint64(32)).astype(np.uint32)], axis=1).copy()
        lut = [interaction_mask(i, o) for o in range(8) for i in range(64)]
        self.lut = np.array([(m & 0xFFFFFFFF, m >> 32) for m in lut], np.uint32)
        self.groups = np.zeros((1 << (2 * (sxz - 2) + sy - 2), 2), np.uint32)
        self.c = EmuScene(self.hdr.ctypes.data, self.cells.ctypes.data, self.voxels.ctypes.data, self.palette.ctypes.data, sxz, sy)


@pytest.fixture(scope="module")
def layout(emu, hash_scene):
    from voxelrt_b200 import capi  # noqa: F401  (HIT_DTYPE)

    L = DeviceLayout(hash_scene)
    emu.emu_build_groups(C.byref(L.c), L.groups.ctypes.data)
    return L


def _trace(emu, L, o, d, wo, flags):
    from voxelrt_b200 import capi

    o = np.ascontiguousarray(o, np.float32)
    d = np.ascontiguousarray(d, np.float32)
    out = np.zeros(len(o), capi.HIT_DTYPE)
    w = (C.c_int32 * 3)(*[int(v) for v in wo])
    emu.emu_trace_glsl(C.byref(L.c), L.groups.ctypes.data, L.lut.ctypes.data, w, o.ctypes.data, d.ctypes.data, flags, len(o), out.ctypes.data)
    return out


def test_group_masks_are_the_sector_occupancy(layout, hash_scene):
    """k_build_groups: bit (x | z<<2 | y<<4) of group (sx>>2, sy>>2, sz>>2) <=> the sector holds bricks (GpuRenderer.cpp:134-142)."""
    want = np.zeros(len(layout.groups), np.uint64)
    for (sx, sy, sz), (mask, _) in hash_scene["sectors"].items():
        if int(mask):
            g = (sx >> 2) | (sz >> 2) << 4 | (sy >> 2) << 8
            want[g] |= np.uint64(1 << ((sx & 3) | (sz & 3) << 2 | (sy & 3) << 4))
    got = layout.groups[:, 0].astype(np.uint64) | (layout.groups[:, 1].astype(np.uint64) << np.uint64(32))
    assert np.array_equal(got, want) and want.any()


@pytest.mark.parametrize("flags", [0, 1, 2, 3])
def test_kernel_source_equals_oracle(emu, layout, hash_oracle, flags):
    for k in range(3):
        wo, o, d = camera_frame_rays(20000, 1200 + k)
        assert_hits_equal(_trace(emu, layout, o, d, wo, flags), hash_oracle.trace_glsl(o, d, wo, flags)[0], f"flags={flags} wo={wo}")
    wo = (96, 64, 96)
    o, d = random_rays(np.random.default_rng(17), 20000, 192, 128, wo)
    assert_hits_equal(_trace(emu, layout, o, d, wo, flags), hash_oracle.trace_glsl(o, d, wo, flags)[0], f"flags={flags} far origins")
    rng = np.random.default_rng(18)
    o = (rng.random((4000, 3)) - 0.5).astype(np.float32)
    d = rng.normal(size=(4000, 3)) * 0.08
    d[:, 1] = -1.0
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    assert_hits_equal(_trace(emu, layout, o, d, (96, 600, 96), flags), hash_oracle.trace_glsl(o, d, (96, 600, 96), flags)[0], f"flags={flags} outside")


def test_kernel_source_equals_oracle_on_special_directions(emu, layout, hash_oracle):
    vals = np.array([0.0, -0.0, 1.0, -1.0, 1e-39, -1e-39, 1e30, np.inf, -np.inf, np.nan, 0.3, -0.7], np.float32)
    d = np.array([(a, b, c) for a in vals for b in vals for c in vals], np.float32)
    o = np.tile(np.array([[0.25, 0.5, 0.75]], np.float32), (len(d), 1))
    o[::7, 0] = np.nan
    for flags in (0, 1, 2):
        assert_hits_equal(_trace(emu, layout, o, d, (90, 70, 90), flags), hash_oracle.trace_glsl(o, d, (90, 70, 90), flags)[0], f"special flags={flags}")
