"""More of the product's kernel source in the CPU tier (tests/native/emu_kernels.cpp): K_hit_query against the reference-generated
golden vectors of VoxelMap::RayCast, the header builders against an independent numpy statement of the layout, the empty-box builder
(k_box_*) against the definition of a box, and — with those boxes in the headers — the traversal WITH macro steps against the oracle:
the jumps must not change a single bit."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import assert_hits_equal
from test_glsl_kernel_on_cpu import DeviceLayout, EmuScene, aligned_zeros
from test_glsl_oracle import camera_frame_rays

NATIVE = Path(__file__).resolve().parent / "native"
GOLD = Path(__file__).resolve().parent / "golden"
OUTSIDE, HASBOX = 0x80000000, 0x40000000


@pytest.fixture(scope="module")
def libs():
    subprocess.run(["make", "-s", "-C", str(NATIVE), "libemu_kernels.so", "libemu_trace.so"], check=True)
    k = C.CDLL(str(NATIVE / "libemu_kernels.so"))
    k.emu_hit_query.argtypes = [C.POINTER(EmuScene), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p]
    k.emu_build_headers.argtypes = [C.POINTER(EmuScene)] + [C.c_void_p] * 4 + [C.c_uint32]
    k.emu_build_boxes.argtypes = [C.POINTER(EmuScene), C.c_void_p]
    t = C.CDLL(str(NATIVE / "libemu_trace.so"))
    t.emu_trace.argtypes = [C.POINTER(EmuScene), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
    for f in (k.emu_hit_query, k.emu_build_headers, k.emu_build_boxes, t.emu_trace):
        f.restype = None
    return k, t


@pytest.fixture(scope="module")
def layout(hash_scene):
    return DeviceLayout(hash_scene)


def test_hit_query_source_reproduces_the_reference_golden_vectors(libs, layout, hash_oracle):
    """K_hit_query (fp64) against VoxelMap::RayCast of the reference itself (tests/golden/ref_hit_query.npz) and the oracle."""
    from voxelrt_b200 import capi

    k, _ = libs
    z = np.load(GOLD / "ref_hit_query.npz")
    o, d = np.ascontiguousarray(z["origin"], np.float64), np.ascontiguousarray(z["dir"], np.float64)
    got = np.zeros(len(o), capi.HITD_DTYPE)
    k.emu_hit_query(C.byref(layout.c), o.ctypes.data, d.ctypes.data, 1024, len(o), got.ctypes.data)
    want = z["hits"]
    hit = want["dist"] >= 0
    assert np.array_equal(got["dist"].view(np.uint64), want["dist"].view(np.uint64)) and hit.sum() > 1000
    for f in ("nx", "ny", "nz", "u", "v"):
        assert np.array_equal(got[f][hit].view(np.uint32), want[f][hit].view(np.uint32)), f
    for f in ("vx", "vy", "vz"):
        assert np.array_equal(got[f][hit], want[f][hit]), f
    rng = np.random.default_rng(9)
    o = rng.uniform(-20, 220, (20000, 3))
    d = rng.normal(size=(20000, 3))
    got = np.zeros(len(o), capi.HITD_DTYPE)
    k.emu_hit_query(C.byref(layout.c), o.ctypes.data, d.ctypes.data, 300, len(o), got.ctypes.data)
    want = hash_oracle.hit_query(o, d, max_iters=300)
    for f in got.dtype.names:
        if f != "_pad":
            a, b = got[f], want[f]
            assert np.array_equal(a.view(np.uint64 if a.dtype.itemsize == 8 else np.uint32), b.view(np.uint64 if b.dtype.itemsize == 8 else np.uint32)), f


def test_header_builders_equal_the_layout_definition(libs, layout, hash_scene):
    """k_init_headers + k_write_headers produce the table the numpy statement of DESIGN.md §4 describes."""
    k, _ = libs
    sxp = (1 << layout.sxz) + 2
    n, guard = len(layout.hdr), 2 * sxp * sxp
    hdr_all = aligned_zeros((n + 2 * guard, 4), np.uint32)
    hdr_all[:] = 0xDEADBEEF
    scene = EmuScene(hdr_all[guard:].ctypes.data, layout.cells.ctypes.data, layout.voxels.ctypes.data, layout.palette.ctypes.data, layout.sxz, layout.sy)
    idx, lo, hi, base, slot = [], [], [], [], 0
    for (sx, sy, sz), (mask, _) in sorted(hash_scene["sectors"].items()):
        mask = int(mask)
        idx.append((sx + 1) + (sz + 1) * sxp + (sy + 1) * sxp * sxp)
        lo.append(mask & 0xFFFFFFFF)
        hi.append(mask >> 32)
        base.append(slot)
        slot += bin(mask).count("1")
    arr = [np.array(a, np.uint32) for a in (idx, lo, hi, base)]
    k.emu_build_headers(C.byref(scene), *[a.ctypes.data for a in arr], len(idx))
    assert np.array_equal(hdr_all, layout.hdr_all)


def test_boxes_are_empty_and_macro_steps_change_nothing(libs, layout, hash_scene, hash_oracle):
    k, t = libs
    from voxelrt_b200 import capi

    sxp, syp = (1 << layout.sxz) + 2, (1 << layout.sy) + 2
    n, guard = len(layout.hdr), 2 * sxp * sxp
    hdr_all = aligned_zeros((n + 2 * guard, 4), np.uint32)
    hdr_all[:] = layout.hdr_all
    hdr = hdr_all[guard : guard + n]
    scene = EmuScene(hdr.ctypes.data, layout.cells.ctypes.data, layout.voxels.ctypes.data, layout.palette.ctypes.data, layout.sxz, layout.sy)
    sat = np.zeros(n, np.uint32)
    k.emu_build_boxes(C.byref(scene), sat.ctypes.data)
    # every box holds only empty in-view sectors and contains its own sector; resident headers are untouched
    grid = hdr.reshape(syp, sxp, sxp, 4)  # [y, z, x]
    occupied = (grid[..., 0] | grid[..., 1]) != 0
    assert np.array_equal(grid[occupied], layout.hdr.reshape(syp, sxp, sxp, 4)[occupied])
    has = (grid[..., 3] & HASBOX) != 0
    assert has.sum() > 1000 and not (has & occupied).any()
    ys, zs, xs = np.nonzero(has)
    pick = np.random.default_rng(1).choice(len(ys), 400, replace=False)
    for y, z, x in zip(ys[pick], zs[pick], xs[pick]):
        lo_c, hi_c = int(grid[y, z, x, 2]), int(grid[y, z, x, 3])
        x0, y0, z0 = lo_c & 255, (lo_c >> 8) & 255, (lo_c >> 16) & 255
        x1, y1, z1 = hi_c & 255, (hi_c >> 8) & 255, (hi_c >> 16) & 255
        assert x0 <= x - 1 <= x1 and y0 <= y - 1 <= y1 and z0 <= z - 1 <= z1
        assert not occupied[y0 + 1 : y1 + 2, z0 + 1 : z1 + 2, x0 + 1 : x1 + 2].any()
        assert max(x1 - x0, y1 - y0, z1 - z0) >= 2
    # the traversal with macro steps over these headers: identical bits to the step-by-step oracle
    def trace(o, d, wo, mode):
        out = np.zeros(len(o), capi.HIT_DTYPE)
        w = (C.c_int32 * 3)(*[int(v) for v in wo])
        t.emu_trace(C.byref(scene), w, np.ascontiguousarray(o, np.float32).ctypes.data, np.ascontiguousarray(d, np.float32).ctypes.data, 0, len(o), out.ctypes.data, mode, None)
        return out

    rng = np.random.default_rng(4)
    cases = [camera_frame_rays(40000, 3000 + i) for i in range(3)]
    for wo in ((1000, 300, 1000), (96, 480, 96), (1900, 100, 30), (5, 5, 2000)):  # far from the terrain: long runs of empty sectors
        o = rng.random((40000, 3)).astype(np.float32)
        d = rng.normal(size=(40000, 3))
        cases.append((wo, o, (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)))
    n_iter_diff = 0
    for wo, o, d in cases:
        want = hash_oracle.trace(o, d, wo)[0]
        got = trace(o, d, wo, 2)
        # a jump over k sectors counts k iterations where the reference needed between 1 and k: the iteration COUNT of
        # a ray that jumped is an upper bound, everything else is bit-identical — the GPU tests compare the same way
        assert_hits_equal(got, want, f"macro steps, wo={wo}", ignore_iters=True)
        assert ((got["flags"] >> 16) >= (want["flags"] >> 16)).all()
        n_iter_diff += int(((got["flags"] >> 16) != (want["flags"] >> 16)).sum())
        assert_hits_equal(trace(o, d, wo, 0), want, f"step by step, wo={wo}")
    assert n_iter_diff < 100  # and it is almost always exact


def test_occupancy_build_kernel_by_warp_replay(libs, layout, hash_scene):
    """K_upload (one warp per dirty brick: copy + the eight 4x4x4 occupancy masks, FlatVoxelStorage::UpdateOccupancy) — its lanes OR their
    partial masks together through xor-shuffles, so the emulator replays the warp once per shuffle call (cuda_host_shim.h: ShflReplay).
    Result against the numpy layout (bit = x + 4z + 16y of every non-empty voxel) and the oracle's orc_build_occupancy, which is pinned
    to the reference's UpdateOccupancy."""
    from oracle import pyoracle

    k, _ = libs
    k.emu_upload_bricks.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    k.emu_upload_bricks.restype = None
    bricks = np.concatenate([np.asarray(b, np.uint8).reshape(-1, 512) for _, (m, b) in sorted(hash_scene["sectors"].items())])
    n = len(bricks)
    assert n * 512 == layout.voxels.size
    rng = np.random.default_rng(12)
    slots = rng.permutation(n).astype(np.uint32)  # bricks land in arbitrary slots
    staging = aligned_zeros(n * 512, np.uint8)
    staging[:] = bricks.reshape(-1)
    voxels = aligned_zeros(n * 512, np.uint8)
    cells = aligned_zeros((n * 8, 2), np.uint32)
    cells[:] = 0xFFFFFFFF
    k.emu_upload_bricks(staging.ctypes.data, slots.ctypes.data, n, voxels.ctypes.data, cells.ctypes.data)
    want_vox = layout.voxels.reshape(n, 512)
    want_cells = layout.cells.reshape(n, 8, 2)
    assert np.array_equal(voxels.reshape(n, 512)[slots], want_vox)
    assert np.array_equal(cells.reshape(n, 8, 2)[slots], want_cells)
    lib = pyoracle.load()
    for i in rng.choice(n, 200, replace=False):
        out = np.zeros(8, np.uint64)
        lib.orc_build_occupancy(bricks[i].ctypes.data, out.ctypes.data)
        got = cells.reshape(n, 8, 2)[slots[i]]
        assert np.array_equal(got[:, 0].astype(np.uint64) | (got[:, 1].astype(np.uint64) << np.uint64(32)), out)


def test_one_bit_sector_table_and_the_occ_loop(libs, layout, hash_oracle):
    """k_build_occ (a ballot per 32 header entries, by warp replay) against its definition, and the OCC form of the step-by-step loop (what the
    frame kernels of big views run for bounce rays) against the oracle."""
    from voxelrt_b200 import capi

    k, t = libs
    k.emu_build_occ.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    k.emu_build_occ.restype = None
    t.emu_set_occ.argtypes = [C.c_void_p]
    t.emu_set_occ.restype = None
    n_all = len(layout.hdr_all)
    occ = np.zeros((n_all + 31) // 32 + 1, np.uint32)
    k.emu_build_occ(layout.hdr_all.ctypes.data, n_all, occ.ctypes.data)
    h = layout.hdr_all
    want_bits = ((h[:, 0] | h[:, 1]) != 0) | ((h[:, 3] & OUTSIDE) != 0)
    got_bits = ((occ[np.arange(n_all) >> 5] >> (np.arange(n_all) & 31).astype(np.uint32)) & 1).astype(bool)
    assert np.array_equal(got_bits, want_bits) and want_bits.any() and not want_bits.all()
    t.emu_set_occ(occ.ctypes.data)
    try:
        rng = np.random.default_rng(31)
        for wo in ((96, 64, 96), (600, 200, 1500), (30, 500, 30)):
            o = rng.uniform(-60, 60, (40000, 3)).astype(np.float32)
            d = rng.normal(size=(40000, 3))
            d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
            out = np.zeros(len(o), capi.HIT_DTYPE)
            w = (C.c_int32 * 3)(*wo)
            t.emu_trace(C.byref(layout.c), w, o.ctypes.data, d.ctypes.data, 0, len(o), out.ctypes.data, 3, None)
            assert_hits_equal(out, hash_oracle.trace(o, d, wo)[0], f"OCC loop wo={wo}")
    finally:
        t.emu_set_occ(None)


def test_brick_relocation_kernel(libs, layout):
    """K_move (slot ranges that shift when a sector gains or loses bricks are moved on the device): voxels and cell masks arrive intact."""
    k, _ = libs
    k.emu_move_bricks.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    k.emu_move_bricks.restype = None
    n = layout.voxels.size // 512
    m = min(n // 2, 3000)
    voxels = aligned_zeros(2 * n * 512, np.uint8)
    cells = aligned_zeros((2 * n * 8, 2), np.uint32)
    voxels[: n * 512] = layout.voxels
    cells[: n * 8] = layout.cells
    rng = np.random.default_rng(5)
    src = rng.choice(n, m, replace=False).astype(np.uint32)
    dst = (n + rng.choice(n, m, replace=False)).astype(np.uint32)  # disjoint from every source
    pairs = np.stack([src, dst], axis=1).copy()
    k.emu_move_bricks(pairs.ctypes.data, m, voxels.ctypes.data, cells.ctypes.data)
    assert np.array_equal(voxels.reshape(-1, 512)[dst], layout.voxels.reshape(-1, 512)[src])
    assert np.array_equal(cells.reshape(-1, 8, 2)[dst], layout.cells.reshape(-1, 8, 2)[src])
    assert np.array_equal(voxels[: n * 512], layout.voxels) and np.array_equal(cells[: n * 8], layout.cells)
