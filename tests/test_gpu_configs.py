"""GPU parity for the remaining BASELINE.json configs (2: bundled model + 1 bounce, 3: large view + 2 bounces,
4: per-frame voxel edits + dirty-brick upload), each against the CPU oracle on the same inputs, bit for bit
(the stated float tolerance of the contract is <= 1 f16 ulp of radiance; the canonical arithmetic gives 0)."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_hits_equal

pytestmark = pytest.mark.gpu


def _frame(cam, w, h, **kw):
    from voxelrt_b200 import capi

    proj, inv, wo, frac = cam.matrices(w, h)
    return capi.make_frame(w, h, inv, proj, wo, frac, **kw)


def _pair(scene, view, shading_inputs=None, capacity=1 << 16):
    from oracle import pyoracle
    from scenes import terrain
    from voxelrt_b200 import capi

    recs = terrain.scene_records(scene)
    ctx = capi.Context(*view, device=0, initial_brick_capacity=capacity)
    orc = pyoracle.OracleMap(*view)
    for m in (ctx, orc):
        m.set_palette(scene["palette"])
        m.sync(recs)
        if shading_inputs is not None:
            (bn, _), (desc, tex, _) = shading_inputs
            m.set_blue_noise(bn)
            m.set_sky(desc, tex)
    return ctx, orc


def _assert_frames_equal(out_g, out_c, what):
    for k in ("albedo", "depth", "irr_rg", "irr_bx"):
        a, b = out_g[k].view(np.uint32), out_c[k].view(np.uint32)
        assert np.array_equal(a, b), f"{what}: {k}: {np.count_nonzero(a != b)} texels differ"


def test_config3_sponza_1080p_one_bounce(shading_inputs):
    """BASELINE configs[2]: the bundled Sponza model (voxelised by scenes/models.py into 1024^3 for the test; the bench
    uses 2048^3), 1920x1080, one bounce of blue-noise diffuse rays, frames 1..3: hit records and radiance bit-exact."""
    from scenes import camera, models

    if not models.sponza_available(1024):
        pytest.skip("scenes/_ref/sponza_1024.dat absent (built by __graft_entry__.build() where the reference assets exist)")
    scene = models.sponza(1024)
    ctx, orc = _pair(scene, (5, 4), shading_inputs, capacity=1 << 17)
    cams = [camera.Camera(pos=(210.3, 80.2, 505.7), yaw=1.5, pitch=-0.15), camera.Camera(pos=(700.1, 300.4, 520.2), yaw=-1.2, pitch=-0.6)]
    for frame_no, cam in ((1, cams[0]), (2, cams[0]), (3, cams[1])):
        out_g, aux_g = ctx.render(_frame(cam, 1920, 1080, bounces=1, frame_no=frame_no), want_aux=True)
        out_c, aux_c, st = orc.render(_frame(cam, 1920, 1080, bounces=1, frame_no=frame_no), want_aux=True)
        assert st.rays > 1920 * 1080 * 3 // 2 and st.hits > 1920 * 1080
        assert_hits_equal(aux_g, aux_c, f"sponza frame {frame_no}", ignore_iters=True)
        _assert_frames_equal(out_g, out_c, f"sponza frame {frame_no}")
    ctx.close()


def test_config4_large_view_two_bounces(shading_inputs):
    """BASELINE configs[3] at test size: a 4096x512x4096-voxel view (128x16x128 sectors: runtime view extent, 4.9 MB of
    sector headers, empty-box corners up to sector 127), terrain over 48x7x48 sectors placed off-origin, 2 bounces."""
    from scenes import camera, terrain

    base = terrain.terrain_fastnoise(48, 7, 48) if terrain.fastnoise_available() else terrain.terrain_hash(16, 4, 16, seed=5)
    shift = (70, 2, 75)  # exercise high sector coordinates
    scene = {"sectors": {(x + shift[0], y + shift[1], z + shift[2]): v for (x, y, z), v in base["sectors"].items()}, "palette": base["palette"]}
    nb = terrain.scene_stats(scene)["bricks"]
    ctx, orc = _pair(scene, (7, 4), shading_inputs, capacity=max(1 << 16, 1 << int(nb + 4096).bit_length()))
    ox, oy, oz = (32 * s for s in shift)
    cams = [camera.Camera(pos=(ox + 700.0, oy + 140.0, oz + 650.0), yaw=1.52, pitch=-0.5), camera.Camera(pos=(ox + 100.5, oy + 230.2, oz + 90.7), yaw=0.7, pitch=-0.35)]
    for i, cam in enumerate(cams):
        out_g, aux_g = ctx.render(_frame(cam, 1280, 720, bounces=2, frame_no=5 + i), want_aux=True)
        out_c, aux_c, st = orc.render(_frame(cam, 1280, 720, bounces=2, frame_no=5 + i), want_aux=True)
        assert st.rays > 1280 * 720 and st.hits > 1280 * 720 // 4
        assert_hits_equal(aux_g, aux_c, f"large view cam {i}", ignore_iters=True)
        _assert_frames_equal(out_g, out_c, f"large view cam {i}")
    ctx.close()


def test_trace_far_origins_shallow_components(shading_inputs):
    """Rays that start ~1000 voxels from the frame origin with one shallow component: tStart = (side - o) / d reaches
    10^4..10^6, its rounding error exceeds the 0.001 step bias and the reference STALLS on cell planes (quirk Q1 on any
    axis, positive directions included).  Empty-box macro jumps must never skip such a stall — this is the case that
    slipped through in round 1 until the per-ray tStart bound was added — so every record must equal the
    oracle's, with macro steps on and off, for explicit rays and for frames whose bounce rays start far from the origin."""
    from scenes import camera, terrain

    base = terrain.terrain_fastnoise(48, 7, 48) if terrain.fastnoise_available() else terrain.terrain_hash(16, 4, 16, seed=5)
    ctx, orc = _pair(base, (6, 4), shading_inputs, capacity=1 << 20)
    rng = np.random.default_rng(21)
    n = 1_000_000
    total_capped = 0
    for wo in ((0, 0, 0), (1500, 200, 1400), (40, 100, 1530)):
        p = np.stack([rng.uniform(0, 1536, n), rng.uniform(0, 224, n), rng.uniform(0, 1536, n)], 1)
        o = (p - np.asarray(wo)).astype(np.float32)
        d = rng.normal(size=(n, 3))
        shallow = rng.integers(0, 3, n)
        d[np.arange(n), shallow] *= 10.0 ** rng.uniform(-4, -1, n)  # one component 10..10^4 times smaller
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d = d.astype(np.float32)
        want, st = orc.trace(o, d, wo)
        total_capped += int(((want["flags"] & 0xFFFF) & capi_flag("CAPPED")).astype(bool).sum())
        for macro in (1, 0):
            ctx.set_option("macro_steps", macro)
            assert_hits_equal(ctx.trace(o, d, wo), want, f"far origins wo={wo} macro={macro}", ignore_iters=bool(macro))
    assert total_capped > 1000  # the stalls this test is about do occur
    ctx.set_option("macro_steps", 1)
    from voxelrt_b200 import capi

    # frames whose bounce rays start up to ~1500 voxels from the frame origin (camera in a corner of the terrain)
    for yaw in (0.8, -2.2):
        cam = camera.Camera(pos=(100.3, 200.2, 90.7), yaw=yaw, pitch=-0.4)
        proj, inv, wo, frac = cam.matrices(640, 360)
        out_g, aux_g = ctx.render(capi.make_frame(640, 360, inv, proj, wo, frac, frame_no=2, bounces=2), want_aux=True)
        out_c, aux_c, _ = orc.render(capi.make_frame(640, 360, inv, proj, wo, frac, frame_no=2, bounces=2), want_aux=True)
        assert_hits_equal(aux_g, aux_c, f"far frame yaw {yaw}", ignore_iters=True)
        _assert_frames_equal(out_g, out_c, f"far frame yaw {yaw}")
    ctx.close()


def _point_source_bursts(ctx, orc, rng, lo, hi, bursts, rays_per_burst, what):
    """Rays that all qualify for macro steps (origins within one voxel of the frame origin, no tiny direction component),
    from many frame origins: the macro path is exercised from every part of the scene, not only from a few cameras."""
    jumps_possible = 0
    for k in range(bursts):
        wo = tuple(int(rng.integers(lo[a], hi[a])) for a in range(3))
        o = rng.uniform(0.0, 1.0, (rays_per_burst, 3)).astype(np.float32)
        d = rng.normal(size=(rays_per_burst, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d = np.where(np.abs(d) < 4.0 / 1024, np.sign(d) * 4.0 / 1024 + (d == 0) * 4.0 / 1024, d).astype(np.float32)
        want, st = orc.trace(o, d, wo)
        ctx.set_option("macro_steps", 1)
        assert_hits_equal(ctx.trace(o, d, wo), want, f"{what}: burst {k} wo={wo}", ignore_iters=True)
        jumps_possible += int(st.rays)
    return jumps_possible


def test_macro_steps_from_everywhere(shading_inputs):
    from scenes import models, terrain

    rng = np.random.default_rng(33)
    base = terrain.terrain_fastnoise(48, 7, 48) if terrain.fastnoise_available() else terrain.terrain_hash(16, 4, 16, seed=5)
    ctx, orc = _pair(base, (6, 4), None, capacity=1 << 20)
    _point_source_bursts(ctx, orc, rng, (0, 60, 0), (1536, 400, 1536), 120, 16384, "terrain")
    ctx.close()
    if models.sponza_available(1024):
        scene = models.sponza(1024)
        ctx, orc = _pair(scene, (5, 4), None, capacity=1 << 17)
        _point_source_bursts(ctx, orc, rng, (0, 0, 250), (1024, 440, 780), 120, 16384, "sponza")
        ctx.close()


def capi_flag(name):
    from voxelrt_b200 import capi

    return getattr(capi, f"VRT_HIT_{name}")


def test_config5_brush_strokes(bench_scene):
    """The reference's brush (capsule r=30: fill, erase, replace strokes) as the edit stream: every frame equals the oracle
    fed with the same records; bricks are allocated by fills and whole sectors change emptiness (empty boxes rebuilt)."""
    from scenes import camera, edits

    frames, world = edits.brush_stroke_frames(bench_scene, 20, seed=5, box=((400, 640), (110, 170), (420, 640)))
    ctx, orc = _pair(bench_scene, (6, 4), capacity=1 << 18)
    cam = camera.Camera()
    for i, recs in enumerate(frames):
        ctx.sync(recs)
        orc.sync(recs)
        if i % 3 == 2 or i == len(frames) - 1:
            out_g, aux_g = ctx.render(_frame(cam, 1280, 720, frame_no=i + 1), want_aux=True)
            out_c, aux_c, _ = orc.render(_frame(cam, 1280, 720, frame_no=i + 1), want_aux=True)
            assert_hits_equal(aux_g, aux_c, f"brush frame {i}", ignore_iters=True)
            assert out_g.tobytes() == out_c.tobytes()
    assert sum(len(r) for r in frames) > 100
    ctx.close()


def test_config5_edit_frames(bench_scene):
    """BASELINE configs[4]: frames of random single-voxel edits (set / clear, bricks allocated on demand), dirty bricks
    uploaded by vrt_sync, then a frame: every frame equals the oracle fed with the same records, and the final
    device state equals a context built from scratch out of the edited world."""
    from scenes import camera, edits, terrain

    frames, world = edits.random_edit_frames(bench_scene, 6, 3000, seed=3)
    ctx, orc = _pair(bench_scene, (6, 4), capacity=1 << 18)
    cam = camera.Camera()
    uploaded = 0
    for i, recs in enumerate(frames):
        ctx.sync(recs)
        orc.sync(recs)
        st = ctx.stats()
        assert st.bricks_uploaded == sum(r[5].shape[0] for r in recs)
        uploaded += st.bricks_uploaded
        out_g, aux_g = ctx.render(_frame(cam, 1280, 720, frame_no=i + 1), want_aux=True)
        out_c, aux_c, _ = orc.render(_frame(cam, 1280, 720, frame_no=i + 1), want_aux=True)
        assert_hits_equal(aux_g, aux_c, f"edit frame {i}", ignore_iters=True)
        assert out_g.tobytes() == out_c.tobytes()
    assert 0 < uploaded < 6 * 3000  # delta upload: far fewer bricks than the scene's 124,705
    fresh_scene = world.to_scene(bench_scene["palette"])
    fresh, _ = _pair(fresh_scene, (6, 4), capacity=1 << 18)
    out_a, _ = ctx.render(_frame(cam, 1280, 720, frame_no=9))
    out_b, _ = fresh.render(_frame(cam, 1280, 720, frame_no=9))
    assert out_a.tobytes() == out_b.tobytes()
    assert ctx.stats().resident_bricks == fresh.stats().resident_bricks == terrain.scene_stats(fresh_scene)["bricks"]
    ctx.close()
    fresh.close()


def test_config5_replicas_stay_identical(bench_scene):
    """BASELINE configs[4] on N GPUs = N replicas of the brickmap, each applying the same edit records (GpuRenderer.cpp:45-167 run once
    per process).  Three contexts stand in for three ranks: after every edit frame their resident state is THE SAME — allocation mask, slot
    base (the arena is deterministic), bricks and cell masks of every edited sector, and equal to the oracle's — and the frame assembled
    from the three ranks' band shares is the frame one context renders alone."""
    import ctypes as C

    from scenes import camera, edits, terrain
    from voxelrt_b200 import capi

    n = 3
    frames, _ = edits.random_edit_frames(bench_scene, 5, 3000, seed=11)
    recs0 = terrain.scene_records(bench_scene)
    ranks = []
    for _ in range(n):
        c = capi.Context(6, 4, device=0, initial_brick_capacity=1 << 18)
        c.set_palette(bench_scene["palette"])
        c.sync(recs0)
        ranks.append(c)
    _, orc = _pair(bench_scene, (6, 4), capacity=1 << 18)
    cam = camera.Camera()
    w, h = 1280, 720
    for i, recs in enumerate(frames):
        for c in ranks:
            c.sync(recs)
        orc.sync(recs)
        stats = [c.stats() for c in ranks]
        assert len({(s.resident_bricks, s.free_ranges, s.resident_sectors, s.bricks_uploaded, s.bricks_relocated) for s in stats}) == 1
        for rec in recs[:: max(1, len(recs) // 60)]:  # a sample of the edited sectors, read back from every replica
            sx, sy, sz = rec[0], rec[1], rec[2]
            states = [c.read_sector(sx, sy, sz) for c in ranks]
            om, ob, oc = orc.read_sector(sx, sy, sz)
            for gm, base, gb, gc in states:
                assert gm == om == states[0][0] and base == states[0][1], (i, sx, sy, sz)
                assert np.array_equal(gb, ob) and np.array_equal(gc, oc), (i, sx, sy, sz)
        host = np.zeros(w * h // 16, capi.TILE_DTYPE)
        for r, c in enumerate(ranks):  # every rank delivers its 8-pixel bands into the one host frame
            f = _frame(cam, w, h, frame_no=i + 1, flags=capi.VRT_FRAME_PART_ROWS, part_index=r, part_count=n)
            c._chk(c.lib.vrt_render(c.h, C.byref(f), host.ctypes.data, None))
        alone, _ = ranks[i % n].render(_frame(cam, w, h, frame_no=i + 1))
        want, _, _ = orc.render(_frame(cam, w, h, frame_no=i + 1))
        assert host.tobytes() == alone.tobytes() == want.tobytes(), f"edit frame {i}"
    for c in ranks:
        c.close()


def test_bricks_beyond_four_gigabytes_of_voxels(hash_scene, hash_oracle, shading_inputs):
    """64-bit brick addressing: with the first 8.5 M slots of the arena reserved, the scene's bricks live where a 10 GB scene's do (byte offsets
    beyond 2^32: 8.4 M slots x 512 B) — upload, read-back, explicit rays, frames in both bounce forms must not care.  (Round 2, found at
    8 GPUs on the 10 GB terrain: the wavefront trace pass computed slot * 512 in 32 bits.)"""
    from scenes import camera, terrain
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    ctx = capi.Context(6, 4, device=0, initial_brick_capacity=1 << 17)
    ctx.set_option("reserve_slots", 8_500_000)
    ctx.set_palette(hash_scene["palette"])
    ctx.sync(terrain.scene_records(hash_scene))
    ctx.set_blue_noise(bn)
    ctx.set_sky(desc, tex)
    hash_oracle.set_blue_noise(bn)
    hash_oracle.set_sky(desc, tex)
    key = sorted(hash_scene["sectors"])[3]
    mask, base, bricks, cells = ctx.read_sector(*key)
    assert base >= 8_500_000 and mask == hash_scene["sectors"][key][0]
    om, ob, oc = hash_oracle.read_sector(*key)
    assert np.array_equal(bricks, ob) and np.array_equal(cells, oc)
    rng = np.random.default_rng(11)
    from conftest import assert_hits_equal, random_rays

    o, d = random_rays(rng, 50_000, 192, 128, (0, 0, 0))
    assert_hits_equal(ctx.trace(o, d, (0, 0, 0)), hash_oracle.trace(o, d, (0, 0, 0))[0], "high slots", ignore_iters=True)
    cam = camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45)
    w, h = 640, 360
    proj, inv, wo, frac = cam.matrices(w, h)
    for bounces in (0, 2):
        want = hash_oracle.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=3, bounces=bounces))[0]
        for wave in (0, 1):
            ctx.set_option("wavefront", wave)
            got, _ = ctx.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=3, bounces=bounces))
            assert got.tobytes() == want.tobytes(), (bounces, wave)
    ctx.close()
