"""The edit tool behind BASELINE configs[4] pinned against the reference's own code: BrushSession::Dispatch (Brush.cpp:10-37) over
VoxelMap::RegionDispatchSIMD (VoxelMap.h:223-263), compiled from where they lie into oracle/_ref/libref_cpu_strict.so
(-ffp-contract=off: the capsule test is a comparison of an fp32 expression with 0, and scenes/brush.cpp writes every fused operation
explicitly).  After every stroke of a session — fill, erase, replace; strokes of zero length; strokes reaching outside the world's
populated part — scenes/edits.py's world and the reference's VoxelMap hold the same voxels in the same bricks (allocation masks included:
bricks created on lookup, emptied bricks and sectors garbage-collected), and the sync records carry exactly VoxelMap::DirtyLocs."""
from __future__ import annotations

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref():
    from oracle import refharness
    from scenes import edits

    if not refharness.available("strict"):
        pytest.skip("oracle/_ref/libref_cpu_strict.so not available on this machine")
    try:
        edits._brushlib()
    except OSError:
        pytest.skip("scenes/_ref/libbrush.so not built")
    return refharness


def _bricks(sectors):
    out = {}
    for key, (mask, bricks) in sectors.items():
        j = 0
        for i in range(64):
            if mask >> i & 1:
                out[key + (i,)] = bricks[j]
                j += 1
    return out


def test_brush_sessions_edit_the_world_like_the_reference(ref, hash_scene):
    from scenes import edits, terrain

    rm = ref.RefMap("strict")
    rm.sync(terrain.scene_records(hash_scene))
    rm.take_dirty()
    world = edits.EditableWorld(hash_scene)
    rng = np.random.default_rng(7)
    box = ((20, 170), (40, 120), (20, 170))
    pos = np.array([rng.integers(lo, hi) for lo, hi in box], np.int64)
    seen = {"fill": 0, "erase": 0, "replace": 0, "gc": 0, "created": 0}
    for f in range(30):
        if f % 10 == 9:
            nxt = pos.copy()  # a click without movement: capsule of zero length (0 / 0 in the projection)
        else:
            delta = rng.normal(size=3)
            nxt = np.array([int(np.clip(pos[a] + int(delta[a] / np.linalg.norm(delta) * 26), box[a][0], box[a][1])) for a in range(3)], np.int64)
        material, action = ((254, "fill"), (0, "replace"), (252, "replace"), (0, "fill"))[(f // 3) % 4]
        radius = (30.0, 12.5, 4.0)[f % 3]
        before = {k: world.alloc_mask(k) for k in world.sectors}
        recs = edits.brush_dispatch(world, pos, nxt, radius, material, action)
        dirty_ref = rm.brush(pos, nxt, radius, action, material)
        got, want = _bricks(world.to_scene(hash_scene["palette"])["sectors"]), _bricks(rm.map_sectors())
        assert set(got) == set(want), f"stroke {f}: allocation"
        assert all(np.array_equal(got[k], want[k]) for k in want), f"stroke {f}: voxels"
        ours = {(r[0], r[1], r[2]): r[4] for r in recs}
        assert ours == {k: v for k, v in dirty_ref.items() if min(k) >= 0}, f"stroke {f}: dirty bricks"
        for r in recs:  # the record contract: payload = dirty & alloc bricks in ascending order; a vanished sector says so
            key = (r[0], r[1], r[2])
            assert r[3] == world.alloc_mask(key) and r[5].shape[0] == bin(r[3] & r[4]).count("1") and r[6] == (key not in world.sectors)
        seen["erase" if material == 0 else action] += 1
        seen["gc"] += sum(1 for k, m in before.items() if m & ~world.alloc_mask(k))
        seen["created"] += sum(1 for k in world.sectors if k not in before)
        pos = nxt
    rm.close()
    assert min(seen.values()) > 0, seen


def test_brush_frames_of_the_bench_replay_on_the_reference(ref, bench_scene):
    """bench.py --edit-mode brush: the first frames of its seeded session, replayed stroke by stroke on the reference's VoxelMap."""
    from scenes import edits, terrain

    frames, world = edits.brush_stroke_frames(bench_scene, 6, seed=1)
    rm = ref.RefMap("strict")
    rm.sync(terrain.scene_records(bench_scene))
    rng = np.random.default_rng(1)  # the session's own walk (brush_stroke_frames)
    box, step = ((120, 640), (100, 170), (120, 640)), 24
    pos = np.array([rng.integers(lo, hi) for lo, hi in box], np.int64)
    for f in range(6):
        delta = rng.normal(size=3)
        delta = (delta / np.linalg.norm(delta) * step).astype(np.int64)
        nxt = np.array([int(np.clip(pos[a] + delta[a], box[a][0], box[a][1])) for a in range(3)], np.int64)
        material, action = ((254, "fill"), (0, "replace"), (252, "replace"))[(f // 8) % 3]
        dirty_ref = rm.brush(pos, nxt, 30.0, action, material)
        assert {(r[0], r[1], r[2]): r[4] for r in frames[f]} == dirty_ref, f
        pos = nxt
    got, want = _bricks(world.to_scene(bench_scene["palette"])["sectors"]), _bricks(rm.map_sectors())
    rm.close()
    assert set(got) == set(want) and all(np.array_equal(got[k], want[k]) for k in want)
