"""Residency parity: after any script of syncs (adds, edits, brick removals, sector removals, arena
growth) the device brickmap equals the oracle's FlatVoxelStorage content, and traversal agrees."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_hits_equal, random_rays

pytestmark = pytest.mark.gpu


def _rand_brick(rng, density):
    b = rng.integers(1, 256, 512, dtype=np.uint8)
    b[rng.random(512) >= density] = 0
    if not b.any():
        b[rng.integers(0, 512)] = 1
    return b


def _compare_all(ctx, orc, n_xz, n_y):
    for sy in range(n_y):
        for sz in range(n_xz):
            for sx in range(n_xz):
                gm, base, gb, gc = ctx.read_sector(sx, sy, sz)
                om, ob, oc = orc.read_sector(sx, sy, sz)
                assert gm == om, (sx, sy, sz)
                assert np.array_equal(gb, ob), (sx, sy, sz)
                assert np.array_equal(gc, oc), (sx, sy, sz)


def test_occupancy_kernel_matches_oracle():
    from oracle import pyoracle
    from voxelrt_b200 import capi

    rng = np.random.default_rng(0)
    ctx = capi.Context(2, 2, device=0)
    bricks = np.stack([_rand_brick(rng, dens) for dens in (0.01, 0.1, 0.5, 0.9, 1.0) for _ in range(12)] + [np.zeros(512, np.uint8)] * 4)
    mask = (1 << 64) - 1
    ctx.sync([(1, 2, 3, mask, mask, bricks)])
    gm, base, gb, gc = ctx.read_sector(1, 2, 3)
    assert gm == mask
    assert np.array_equal(gb, bricks)
    for i in range(64):
        assert np.array_equal(gc[i], pyoracle.build_occupancy(bricks[i])), i
    ctx.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_random_edit_script(seed):
    """Config-5 style workload: frames of random brick edits / allocations / removals with
    dirty-brick delta upload; device state == oracle state after every frame."""
    from oracle import pyoracle
    from voxelrt_b200 import capi

    rng = np.random.default_rng(seed)
    n_xz, n_y = 4, 2
    ctx = capi.Context(2, 1, device=0, initial_brick_capacity=64)  # tiny arena: forces growth + relocation
    orc = pyoracle.OracleMap(2, 1)
    pal = rng.integers(0, 1 << 40, 256, dtype=np.uint64)
    ctx.set_palette(pal)
    orc.set_palette(pal)
    world = {}  # (sx,sy,sz) -> {brick: bytes}
    for frame in range(24):
        recs = []
        for _ in range(rng.integers(1, 9)):
            key = (int(rng.integers(0, n_xz)), int(rng.integers(0, n_y)), int(rng.integers(0, n_xz)))
            if any(r[:3] == key for r in recs):
                continue
            sec = world.setdefault(key, {})
            op = rng.random()
            dirty = 0
            if op < 0.1 and sec:  # sector deleted
                world.pop(key)
                recs.append((*key, 0, (1 << 64) - 1, None, True))
                continue
            for _ in range(rng.integers(1, 12)):
                b = int(rng.integers(0, 64))
                r = rng.random()
                if r < 0.25 and b in sec:
                    del sec[b]  # brick freed (RegionDispatchSIMD GC, VoxelMap.h:254-262)
                    dirty |= 1 << b
                else:
                    sec[b] = _rand_brick(rng, rng.choice([0.02, 0.3, 0.8]))
                    dirty |= 1 << b
            alloc = sum(1 << b for b in sec)
            payload = [sec[b] for b in sorted(sec) if dirty >> b & 1]
            recs.append((*key, alloc, dirty, np.stack(payload) if payload else None))
        ctx.sync(recs)
        orc.sync(recs)
        st = ctx.stats()
        assert st.resident_bricks == sum(len(s) for s in world.values())
        if frame % 6 == 5:
            _compare_all(ctx, orc, n_xz, n_y)
            o, d = random_rays(rng, 20_000, 128, 64, (0, 0, 0))
            assert_hits_equal(ctx.trace(o, d, (0, 0, 0)), orc.trace(o, d, (0, 0, 0))[0], f"frame {frame}", ignore_iters=True)
    _compare_all(ctx, orc, n_xz, n_y)
    assert ctx.stats().brick_capacity > 64
    ctx.close()


def test_out_of_view_sectors_are_skipped():
    from voxelrt_b200 import capi

    ctx = capi.Context(2, 1, device=0)
    b = np.full((1, 512), 7, np.uint8)
    ctx.sync([(4, 0, 0, 1, 1, b), (0, 2, 0, 1, 1, b), (-1, 0, 0, 1, 1, b), (3, 1, 3, 1, 1, b)])
    assert ctx.stats().resident_bricks == 1  # CheckInBounds -> continue (CpuRenderer.cpp:40)
    ctx.close()


def test_delta_upload_moves_only_dirty_bricks(hash_scene):
    """Dirty-brick delta upload: editing one brick of a resident scene uploads one brick."""
    from conftest import ctx_for

    ctx = ctx_for(hash_scene)
    full = ctx.stats().bytes_uploaded
    (sx, sy, sz), (mask, bricks) = next(iter(sorted(hash_scene["sectors"].items())))
    b0 = int(mask & -mask).bit_length() - 1
    edited = bricks[0].copy()
    edited[:8] = 9
    ctx.sync([(sx, sy, sz, mask, 1 << b0, edited[None])])
    st = ctx.stats()
    assert st.bricks_uploaded == 1 and st.bricks_relocated == 0
    assert st.bytes_uploaded < 1024 < full
    gm, base, gb, gc = ctx.read_sector(sx, sy, sz)
    assert np.array_equal(gb[b0], edited)
    ctx.close()


def test_allocated_but_not_dirty_bricks_are_empty():
    """A brick that enters alloc_mask without being dirty (VoxelMap::GetBrick creates bricks on lookup, quirk Q5) must be resident
    as an EMPTY brick even when its slot is recycled from a freed sector (recycled slots held stale voxels)."""
    from oracle import pyoracle
    from voxelrt_b200 import capi

    rng = np.random.default_rng(5)
    ctx = capi.Context(2, 1, device=0, initial_brick_capacity=64)
    orc = pyoracle.OracleMap(2, 1)
    pal = rng.integers(0, 1 << 40, 256, dtype=np.uint64)
    ctx.set_palette(pal)
    orc.set_palette(pal)
    full = (1 << 64) - 1
    solid = np.stack([_rand_brick(rng, 0.9) for _ in range(64)])
    for m in (ctx, orc):
        m.sync([(0, 0, 0, full, full, solid)])           # fills the whole 64-slot arena with solid bricks
        m.sync([(0, 0, 0, 0, full, None, True)])         # ... and frees it again: every slot now holds stale voxels
    dirty = 0b11111
    alloc = 0b1111111111 | (1 << 40)
    payload = np.stack([_rand_brick(rng, 0.5) for _ in range(5)])
    for m in (ctx, orc):
        m.sync([(1, 0, 2, alloc, dirty, payload)])       # bricks 5-9 and 40: allocated, never written
    _compare_all(ctx, orc, 4, 2)
    gm, base, gb, gc = ctx.read_sector(1, 0, 2)
    assert gm == alloc and not gb[5:10].any() and not gb[40].any() and not gc[5:10].any() and not gc[40].any()
    # same through the relocation path: the sector gains a low brick (everything moves) plus one more empty brick
    alloc2 = alloc | (1 << 3 << 8) | (1 << 50)
    one = _rand_brick(rng, 0.7)[None]
    for m in (ctx, orc):
        m.sync([(1, 0, 2, alloc2, 1 << 11, one)])
    _compare_all(ctx, orc, 4, 2)
    o, d = random_rays(rng, 20_000, 128, 64, (0, 0, 0))
    assert_hits_equal(ctx.trace(o, d, (0, 0, 0)), orc.trace(o, d, (0, 0, 0))[0], "fresh bricks", ignore_iters=True)
    ctx.close()


def test_sync_is_transactional_on_bad_records(hash_scene):
    """A record that fails validation in the MIDDLE of a batch, or a duplicate sector, leaves residency and rendering unchanged
    (VERDICT r1 weak #8 / ADVICE: earlier records of the batch used to mutate the host mirror before the error return)."""
    import ctypes as C

    from conftest import ctx_for
    from voxelrt_b200 import capi

    ctx = ctx_for(hash_scene)
    rng = np.random.default_rng(9)
    keys = sorted(hash_scene["sectors"])[:3]
    before = [ctx.read_sector(*k) for k in keys]
    stats0 = ctx.stats()
    o, d = random_rays(rng, 20_000, 128, 64, (0, 0, 0))
    hits0 = ctx.trace(o, d, (0, 0, 0))
    full = (1 << 64) - 1
    solid = np.full((64, 512), 3, np.uint8)
    # record 0 would re-allocate its sector to 64 bricks (relocation + growth); record 1 claims dirty bricks but has no payload
    arr, keep, n = capi.make_records([(*keys[0], full, full, solid), (*keys[1], full, full, None), (*keys[2], full, full, solid)])
    assert ctx.lib.vrt_sync(ctx.h, n, arr) == capi.VRT_ERR_INVALID
    assert b"nothing was changed" in ctx.lib.vrt_last_error(ctx.h)
    # the same sector twice in one call
    arr2, keep2, n2 = capi.make_records([(*keys[0], full, full, solid), (*keys[2], 1, 1, solid[:1]), (*keys[0], 1, 1, solid[:1])])
    assert ctx.lib.vrt_sync(ctx.h, n2, arr2) == capi.VRT_ERR_INVALID
    after = [ctx.read_sector(*k) for k in keys]
    for (m0, b0, v0, c0), (m1, b1, v1, c1) in zip(before, after):
        assert m0 == m1 and b0 == b1 and np.array_equal(v0, v1) and np.array_equal(c0, c1)
    st = ctx.stats()
    assert (st.resident_bricks, st.resident_sectors, st.free_ranges) == (stats0.resident_bricks, stats0.resident_sectors, stats0.free_ranges)
    assert_hits_equal(ctx.trace(o, d, (0, 0, 0)), hits0, "after rejected syncs")
    # and the context still works: the valid part of the batch goes through afterwards
    ctx.sync([(*keys[0], full, full, solid), (*keys[2], full, full, solid)])
    gm, _, gb, _ = ctx.read_sector(*keys[0])
    assert gm == full and np.array_equal(gb, solid)
    ctx.close()


@pytest.mark.parametrize("threads", [0, 1, 3])
def test_staging_gather_threads_deliver_the_same_bricks(threads):
    """vrt_sync shares the staging gather of a batch of >= 8192 bricks out over a pool of host threads that lives with the context
    (option gather_threads: 0 = min(8, cores), 1 = calling thread only): every split delivers every brick to its slot, call after call
    (the pool sleeps in between), also after the thread count changes."""
    from voxelrt_b200 import capi

    rng = np.random.default_rng(5)
    ctx = capi.Context(4, 2, device=0)
    ctx.set_option("gather_threads", threads)
    full = (1 << 64) - 1
    state = {}
    for rnd in range(4):
        recs = []
        for sx, sz, sy in [(x, z, y) for y in range(2) for z in range(8) for x in range(8)]:
            bricks = rng.integers(0, 256, (64, 512), dtype=np.uint8)
            bricks[:, 0] = 1 + (rnd & 1)  # never all-empty
            recs.append((sx, sy, sz, full, full, bricks))
            state[(sx, sy, sz)] = bricks
        ctx.sync(recs)  # 128 sectors x 64 bricks = 8192 bricks in one batch
        assert ctx.stats().bricks_uploaded == 8192
        for (sx, sy, sz), want in state.items():
            gm, base, gb, gc = ctx.read_sector(sx, sy, sz)
            assert gm == full and np.array_equal(gb, want), (rnd, sx, sy, sz)
        if rnd == 1:
            ctx.set_option("gather_threads", 2 if threads != 2 else 4)  # stops the pool; the next big batch starts a new one
    ctx.close()
