"""CPU tests of the scene INPUT generators that feed configs 3-5 (voxeliser, palette quantiser, edit stream)."""
from __future__ import annotations

import numpy as np
import pytest


def _voxelize(tris, size=32):
    import ctypes as C

    from scenes import models

    if not models.VOX_LIB.exists():
        pytest.skip("scenes/_ref/libvoxelizer.so not built")
    lib = models._voxlib()
    g = lib.vox_create(size // 8, size // 8, size // 8)
    vt = np.ascontiguousarray(tris, np.float32)
    uv = np.zeros((vt.shape[0], 3, 2), np.float32)
    tex = np.full(vt.shape[0], -1, np.int32)
    ptrs = (C.c_void_p * 1)()
    one = np.ones(1, np.int32)
    pal = np.zeros((1, 3), np.uint8)
    lib.vox_triangles(g, vt.shape[0], vt.ctypes.data, uv.ctypes.data, tex.ctypes.data, ptrs, one.ctypes.data, one.ctypes.data, pal.ctypes.data, 1, 7)
    n = lib.vox_brick_count(g)
    bricks = np.zeros((n, 512), np.uint8)
    coords = np.zeros((n, 3), np.int32)
    lib.vox_export(g, bricks.ctypes.data, coords.ctypes.data)
    lib.vox_destroy(g)
    vol = np.zeros((size, size, size), np.uint8)  # [y, z, x]
    for c, b in zip(coords, bricks):
        vol[c[1] * 8 : c[1] * 8 + 8, c[2] * 8 : c[2] * 8 + 8, c[0] * 8 : c[0] * 8 + 8] = b.reshape(8, 8, 8)
    return vol


def test_voxelizer_axis_aligned_quad():
    # a quad in the plane z = 3.5 over [2,10] x [4,9] (x, y): exactly the voxels with z = 3 whose cell touches it
    a, b, c, d = (2.0, 4.0, 3.5), (10.0, 4.0, 3.5), (10.0, 9.0, 3.5), (2.0, 9.0, 3.5)
    vol = _voxelize([[a, b, c], [a, c, d]])
    ys, zs, xs = np.nonzero(vol)
    assert set(zs.tolist()) == {3}
    assert xs.min() <= 2 and xs.max() >= 9 and ys.min() <= 4 and ys.max() >= 8
    assert (vol[4:9, 3, 2:10] == 7).all()  # the interior is watertight
    assert xs.min() >= 1 and xs.max() <= 10 and ys.min() >= 3 and ys.max() <= 9  # conservative by at most one cell


def test_voxelizer_covers_every_sampled_surface_point():
    rng = np.random.default_rng(5)
    tris = rng.uniform(3, 28, (40, 3, 3)).astype(np.float32)
    vol = _voxelize(tris)
    # conservative voxelisation: every voxel that contains a point of a triangle is set ...
    w = rng.dirichlet((1, 1, 1), 4000).astype(np.float32)
    for t in tris:
        p = w @ t
        cell = np.floor(p).astype(int)
        assert (vol[cell[:, 1], cell[:, 2], cell[:, 0]] == 7).all()
    # ... and every set voxel's cube is within reach of some triangle's plane and bounding box
    ys, zs, xs = np.nonzero(vol)
    centers = np.stack([xs, ys, zs], 1) + 0.5
    ok = np.zeros(len(centers), bool)
    for t in tris:
        n = np.cross(t[1] - t[0], t[2] - t[0])
        n /= np.linalg.norm(n)
        near_plane = np.abs((centers - t[0]) @ n) <= 0.5 * np.abs(n).sum() + 1e-4
        in_box = ((centers + 0.5 >= t.min(0) - 1e-4) & (centers - 0.5 <= t.max(0) + 1e-4)).all(1)
        ok |= near_plane & in_box
    assert ok.all()


def test_octree_palette_and_nearest_index():
    from scenes import models

    rng = np.random.default_rng(2)
    cols = np.concatenate([rng.integers(0, 256, (5000, 3)), np.tile([[200, 30, 30]], (3000, 1)), np.tile([[10, 10, 240]], (10, 1))]).astype(np.uint8)
    q = models.OctreePalette()
    q.add_colors(cols[:4000])
    q.add_colors(cols[4000:])
    pal = q.build(240)
    assert 200 <= pal.shape[0] <= 240
    idx = models.nearest_palette_index(pal, cols)
    d = np.abs(cols[:, None, :].astype(int) - pal[None].astype(int)).sum(2)
    assert np.array_equal(idx, d.argmin(1))
    assert d[np.arange(len(cols)), idx].mean() < 40  # the palette actually represents the input
    # a heavily populated colour keeps (nearly) its own entry
    assert d[5000, idx[5000]] <= 6


def test_edit_stream_records_reproduce_the_world(hash_scene):
    """Feeding the per-frame dirty records to a map incrementally gives the same state as syncing the edited world from
    scratch (the VrtDirtySector contract: alloc mask + dirty mask + payload of dirty & alloc in ascending brick order)."""
    from oracle import pyoracle
    from scenes import edits, terrain

    frames, world = edits.random_edit_frames(hash_scene, 5, 800, seed=9, box=((0, 192), (0, 128), (0, 192)))
    inc = pyoracle.OracleMap(6, 4)
    inc.sync(terrain.scene_records(hash_scene))
    for recs in frames:
        assert all(r[5].shape[0] == bin(r[3] & r[4]).count("1") for r in recs)
        inc.sync(recs)
    fresh = pyoracle.OracleMap(6, 4)
    final = world.to_scene(hash_scene["palette"])
    fresh.sync(terrain.scene_records(final))
    assert terrain.scene_stats(final)["bricks"] > terrain.scene_stats(hash_scene)["bricks"]  # edits allocated bricks
    for key in sorted(final["sectors"]):
        a, b = inc.read_sector(*key), fresh.read_sector(*key)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), key


def test_cvox_reads_the_reference_written_golden_file_and_round_trips(tmp_path):
    """tests/golden/ref_cvox_small.dat was written by the reference's own VoxelMap::Serialize (make_golden_cvox.py):
    scenes/cvox.py reads it back to the content that went in, and what it writes it reads again (the live, both-direction
    pin against the reference is tests/test_ref_pin.py::test_cvox_files_are_wire_compatible_with_the_reference)."""
    from pathlib import Path

    from scenes import cvox, terrain

    g = Path(__file__).parent / "golden"
    want = np.load(g / "ref_cvox_small.npz")
    got = cvox.load_cvox(g / "ref_cvox_small.dat")
    keys = [tuple(int(v) for v in k) for k in want["keys"]]
    assert sorted(got["sectors"]) == keys
    off = 0
    for k, m in zip(keys, want["masks"].tolist()):
        n = bin(m).count("1")
        assert got["sectors"][k][0] == m and np.array_equal(got["sectors"][k][1], want["bricks"][off : off + n]), k
        off += n
    for a, b in zip(got["materials"], want["materials"]):
        assert a[:4] == tuple(int(v) for v in b[:4]) and a[4] == np.float32(b[4])
    out = tmp_path / "again.dat"
    cvox.save_cvox(got, out)
    again = cvox.load_cvox(out)
    assert terrain.scene_digest(again) == terrain.scene_digest(got) and again["materials"] == got["materials"]
    assert cvox.sector_pos(cvox.sector_index(-5, -1, 7)) == (-5, -1, 7) and cvox.sector_pos(cvox.sector_index(2047, 127, -2048)) == (2047, 127, -2048)


def test_brush_capsule_matches_the_distance_function(hash_scene):
    """scenes/edits.brush_dispatch against a direct evaluation of the brush's capsule distance function (Brush.cpp:4-8,19-30)
    over the whole bounding box: fill sets exactly the voxels whose centre is inside, replace only the non-empty ones,
    erase clears them, and the records rebuild the world."""
    from oracle import pyoracle
    from scenes import edits, terrain

    world = edits.EditableWorld(hash_scene)
    a, b, r = (40, 70, 50), (75, 82, 61), 12.0

    def dense(w, lo, hi):
        out = np.zeros((hi[1] - lo[1], hi[2] - lo[2], hi[0] - lo[0]), np.uint8)
        for (sx, sy, sz), d in w.sectors.items():
            for bi, vox in d.items():
                x0, y0, z0 = sx * 32 + (bi & 3) * 8, sy * 32 + (bi >> 4) * 8, sz * 32 + ((bi >> 2) & 3) * 8
                if x0 + 8 <= lo[0] or x0 >= hi[0] or y0 + 8 <= lo[1] or y0 >= hi[1] or z0 + 8 <= lo[2] or z0 >= hi[2]:
                    continue
                for y in range(8):
                    for z in range(8):
                        for x in range(8):
                            X, Y, Z = x0 + x, y0 + y, z0 + z
                            if lo[0] <= X < hi[0] and lo[1] <= Y < hi[1] and lo[2] <= Z < hi[2]:
                                out[Y - lo[1], Z - lo[2], X - lo[0]] = vox[x | (z << 3) | (y << 6)]
        return out

    lo, hi = (24, 56, 32), (96, 96, 80)
    before = dense(world, lo, hi)
    ys, zs, xs = np.meshgrid(np.arange(lo[1], hi[1]), np.arange(lo[2], hi[2]), np.arange(lo[0], hi[0]), indexing="ij")
    # the distance function in float64; voxels within 1e-2 of the surface are left to the bit-exact pin (tests/test_ref_brush_pin.py:
    # the brush works in fp32 with an approximate square root)
    p = np.stack([xs + 0.5, ys + 0.5, zs + 0.5], axis=-1).astype(np.float64)
    pa, ba = p - np.array(a, np.float64), np.array(b, np.float64) - np.array(a, np.float64)
    hh = np.clip((pa @ ba) / (ba @ ba), 0.0, 1.0)
    dist = np.linalg.norm(pa - hh[..., None] * ba, axis=-1) - r
    inside, sure = dist < 0, np.abs(dist) > 1e-2
    assert 20_000 < inside.sum() and sure.mean() > 0.99
    recs_all = []
    recs_all.append(edits.brush_dispatch(world, a, b, r, 253, "fill"))
    after_fill = dense(world, lo, hi)
    assert np.array_equal(after_fill[sure], np.where(inside, 253, before)[sure])
    recs_all.append(edits.brush_dispatch(world, a, b, r, 0, "replace"))
    assert np.array_equal(dense(world, lo, hi)[sure], np.where(inside, 0, before)[sure])
    world2 = edits.EditableWorld(hash_scene)
    edits.brush_dispatch(world2, a, b, r, 7, "replace")
    assert np.array_equal(dense(world2, lo, hi)[sure], np.where(inside & (before != 0), 7, before)[sure])
    # the records rebuild the edited world
    inc = pyoracle.OracleMap(6, 4)
    inc.sync(terrain.scene_records(hash_scene))
    for recs in recs_all:
        inc.sync(recs)
    fresh = pyoracle.OracleMap(6, 4)
    final = world.to_scene(hash_scene["palette"])
    fresh.sync(terrain.scene_records(final))
    for key in sorted(final["sectors"]):
        x, y = inc.read_sector(*key), fresh.read_sector(*key)
        assert x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]), key
