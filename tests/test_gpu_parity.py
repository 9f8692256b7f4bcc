"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit for bit.

Contract (BASELINE.json north_star): hit voxel coordinates, face normals and material ids are
bit-exact; here the canonical arithmetic makes distances, positions, UVs and the
packed G-buffer bit-exact as well, so every comparison below is on raw bytes.
"""
from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_hits_equal, ctx_for, random_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1], ids=["stepwise", "macro"])
def macro(request, hash_ctx, bench_ctx):
    """Runs a test with the empty-box macro steps off and on.  With macro steps the iteration count in
    VrtHit.flags is not the reference's (it is an upper bound); everything else must still be bit-exact."""
    for c in (hash_ctx, bench_ctx):
        c.set_option("macro_steps", request.param)
    yield bool(request.param)
    for c in (hash_ctx, bench_ctx):
        c.set_option("macro_steps", 1)


def _frame(cam, w, h, **kw):
    from voxelrt_b200 import capi

    proj, inv, wo, frac = cam.matrices(w, h)
    return capi.make_frame(w, h, inv, proj, wo, frac, **kw)


# ---------------------------------------------------------------------------------------------
# explicit rays: vrt_trace == RayCast (CpuRenderer.cpp:172-224)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_trace_random_rays(hash_ctx, hash_oracle, seed, macro):
    rng = np.random.default_rng(seed)
    wo = (rng.integers(0, 192), rng.integers(0, 128), rng.integers(0, 192))
    o, d = random_rays(rng, 200_000, 192, 128, wo)
    got = hash_ctx.trace(o, d, wo)
    want, st = hash_oracle.trace(o, d, wo)
    assert st.hits > 20_000
    assert_hits_equal(got, want, f"seed {seed}", ignore_iters=macro)


def test_trace_edge_cases(hash_ctx, hash_oracle, macro):
    """Zero / negative-zero / denormal / inf / NaN direction components, axis-aligned rays, rays
    that start inside solid voxels, outside the grid, on integer coordinates, huge origins."""
    rng = np.random.default_rng(11)
    n = 60_000
    wo = (64, 40, 64)
    o, d = random_rays(rng, n, 192, 128, wo)
    specials = np.array([0.0, -0.0, 1e-40, -1e-40, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-30, 3e38], np.float32)
    for k in range(0, n // 2):
        axis = rng.integers(0, 3)
        d[k, axis] = specials[rng.integers(0, specials.size)]
        if k % 3 == 0:
            d[k, (axis + 1) % 3] = specials[rng.integers(0, specials.size)]
        if k % 7 == 0:
            d[k] = specials[rng.integers(0, specials.size, 3)]
    o[n // 2 : n // 2 + 5000] = np.round(o[n // 2 : n // 2 + 5000])  # integer coordinates
    o[n // 2 + 5000 : n // 2 + 6000] *= np.float32(1e6)  # far outside
    o[n // 2 + 6000 : n // 2 + 6100] = np.float32(np.nan)
    o[n // 2 + 6100 : n // 2 + 6200] = np.float32(3e9)
    got = hash_ctx.trace(o, d, wo)
    want, _ = hash_oracle.trace(o, d, wo)
    assert_hits_equal(got, want, "edge cases", ignore_iters=macro)


@pytest.mark.parametrize("max_iters", [1, 2, 7, 128, 1000])
def test_trace_iteration_cap(hash_ctx, hash_oracle, max_iters, macro):
    rng = np.random.default_rng(5)
    wo = (96, 64, 96)
    o, d = random_rays(rng, 50_000, 192, 128, wo)
    got = hash_ctx.trace(o, d, wo, max_iters=max_iters)
    want, _ = hash_oracle.trace(o, d, wo, max_iters=max_iters)
    assert_hits_equal(got, want, f"max_iters {max_iters}", ignore_iters=macro)


def test_trace_empty_and_ragged(hash_ctx, hash_oracle):
    from voxelrt_b200 import capi

    assert hash_ctx.trace(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), (0, 0, 0)).shape == (0,)
    rng = np.random.default_rng(9)
    for n in (1, 31, 33, 127, 129, 1000):
        o, d = random_rays(rng, n, 192, 128, (0, 0, 0))
        assert_hits_equal(hash_ctx.trace(o, d, (0, 0, 0)), hash_oracle.trace(o, d, (0, 0, 0))[0], f"n={n}", ignore_iters=True)
    with pytest.raises(capi.VrtError):
        hash_ctx._chk(hash_ctx.lib.vrt_trace(hash_ctx.h, 4, None, None, None, 0, None))


def test_metrics_match_oracle_counters(hash_ctx, hash_oracle):
    """The device counters behind roofline.achieved are the oracle's I_s / I_c / H."""
    rng = np.random.default_rng(21)
    o, d = random_rays(rng, 100_000, 192, 128, (0, 0, 0))
    hash_ctx.set_option("metrics", 1)
    try:
        hash_ctx.trace(o, d, (0, 0, 0))
        m = hash_ctx.metrics()
    finally:
        hash_ctx.set_option("metrics", 0)
    _, st = hash_oracle.trace(o, d, (0, 0, 0))
    assert (m.rays, m.sector_fetches, m.cell_fetches, m.hits, m.capped) == (
        st.rays,
        st.sector_fetches,
        st.cell_fetches,
        st.hits,
        st.capped,
    )


# ---------------------------------------------------------------------------------------------
# frames: vrt_render == RenderRow (CpuRenderer.cpp:326-402)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [(256, 144), (260, 148), (36, 4)])
@pytest.mark.parametrize("linear", [False, True])
def test_render_primary_small(hash_ctx, hash_oracle, size, linear, macro):
    from scenes import camera
    from voxelrt_b200 import capi

    w, h = size
    cam = camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45)
    flags = capi.VRT_FRAME_LINEAR_OUTPUT if linear else 0
    out_g, aux_g = hash_ctx.render(_frame(cam, w, h, flags=flags), want_aux=True)
    out_c, aux_c, _ = hash_oracle.render(_frame(cam, w, h, flags=flags), want_aux=True)
    assert_hits_equal(aux_g, aux_c, "primary aux", ignore_iters=macro)
    assert out_g.tobytes() == out_c.tobytes()


def test_render_rejects_bad_sizes(hash_ctx):
    from scenes import camera
    from voxelrt_b200 import capi

    cam = camera.Camera()
    for w, h in ((0, 16), (18, 16), (16, 6)):
        with pytest.raises(capi.VrtError):
            hash_ctx.render(_frame(cam, w, h))


def test_render_config1_720p(bench_ctx, bench_oracle, macro):
    """BASELINE.json configs[0]: reference camera, 1280x720 primary rays, every hit record bit-exact."""
    from scenes import camera

    cam = camera.Camera()
    out_g, aux_g = bench_ctx.render(_frame(cam, 1280, 720), want_aux=True)
    out_c, aux_c, st = bench_oracle.render(_frame(cam, 1280, 720), want_aux=True)
    assert st.rays == 1280 * 720
    assert_hits_equal(aux_g, aux_c, "config 1", ignore_iters=macro)
    assert out_g.tobytes() == out_c.tobytes()
    # without the aux records the host-buffer call takes the band-pipelined path (8 launches, copies overlapped)
    out_b, _ = bench_ctx.render(_frame(cam, 1280, 720))
    assert out_b.tobytes() == out_c.tobytes()
    assert bench_ctx.stats().last_launches == 8


def test_render_config2_4k_full(bench_ctx, bench_oracle, macro):
    """BASELINE.json configs[1] at FULL size: all 8,294,400 primary rays {hit, voxel, normal,
    material} (in fact the whole record) bit-exact against the oracle."""
    from scenes import camera

    cam = camera.Camera()
    out_g, aux_g = bench_ctx.render(_frame(cam, 3840, 2160), want_aux=True)
    out_c, aux_c, st = bench_oracle.render(_frame(cam, 3840, 2160), want_aux=True)
    assert st.rays == 3840 * 2160
    assert_hits_equal(aux_g, aux_c, "config 2", ignore_iters=macro)
    assert out_g.tobytes() == out_c.tobytes()


def test_rcp_rn_normal_matches_ieee(hash_ctx):
    """The unguarded reciprocal (MUFU.RCP + one FMA Newton step) and square root (MUFU.RSQ + one Newton step) are the
    correctly rounded 1/x and sqrt(x) wherever the kernels use them: exhaustive over all binary32 patterns with
    2^-100 <= |x| <= 2^100 against rcp.rn / sqrt.rn on the device (1.68 G patterns, both signs for 1/x), and a sample
    against the host's IEEE division and square root."""
    import ctypes as C

    lo = int(np.float32(2.0**-100).view(np.uint32))
    hi = int(np.float32(2.0**100).view(np.uint32))
    bad = C.c_uint64(123)
    n = 1 << 20
    sample = np.zeros(2 * n, np.float32)
    fn = hash_ctx.lib.vrt_debug_rcp_check
    fn.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint32]
    fn.restype = C.c_int
    assert fn(hash_ctx.h, lo, hi, C.byref(bad), None, 0) == 0
    assert bad.value == 0
    # host cross-check on 2^20 patterns around 1.0 (and their negations)
    lo1 = int(np.float32(0.75).view(np.uint32))
    assert fn(hash_ctx.h, lo1, lo1 + n - 1, C.byref(bad), sample.ctypes.data, n) == 0
    x = (np.arange(n, dtype=np.uint32) + np.uint32(lo1)).view(np.float32)
    want = np.concatenate([np.float32(1.0) / x, np.float32(1.0) / -x])
    assert np.array_equal(sample.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("bounces", [0, 1])
def test_render_persistent_kernel_equals_grid_kernel(bench_ctx, bench_oracle, shading_inputs, bounces):
    """The persistent form of the frame kernel (warps pull tiles from a ticket counter) must produce the very bytes of the
    one-CTA-per-4-tiles form, launch after launch (the counter is never reset), for full frames, tile partitions and
    the band-pipelined host-buffer path, including grids larger and smaller than the tile count."""
    from scenes import camera

    (bn, _), (sky_desc, sky_texels, _) = shading_inputs
    bench_ctx.set_blue_noise(bn)
    bench_ctx.set_sky(sky_desc, sky_texels)
    cam = camera.Camera()
    sizes = [(1280, 720), (36, 4), (260, 148), (1920, 1080)]
    try:
        for w, h in sizes:
            bench_ctx.set_option("persistent", 0)
            want, _ = bench_ctx.render(_frame(cam, w, h, bounces=bounces, frame_no=3))
            for mode in (1, 1, 3):
                bench_ctx.set_option("persistent", mode)
                got, _ = bench_ctx.render(_frame(cam, w, h, bounces=bounces, frame_no=3))
                assert got.tobytes() == want.tobytes(), (w, h, mode)
                got_aux, _ = bench_ctx.render(_frame(cam, w, h, bounces=bounces, frame_no=3), want_aux=True)
                assert got_aux.tobytes() == want.tobytes(), (w, h, mode, "aux path")
        if bounces == 0:
            out_c, _, _ = bench_oracle.render(_frame(cam, 1280, 720), want_aux=False)
            bench_ctx.set_option("persistent", 1)
            got, _ = bench_ctx.render(_frame(cam, 1280, 720))
            assert got.tobytes() == out_c.tobytes()
    finally:
        bench_ctx.set_option("persistent", 0)


@pytest.mark.parametrize("i", range(4))
def test_render_orbit_cameras(bench_ctx, bench_oracle, i, macro):
    from scenes import camera

    cam = camera.orbit_cameras(4, seed=1)[i]
    out_g, aux_g = bench_ctx.render(_frame(cam, 640, 360), want_aux=True)
    out_c, aux_c, _ = bench_oracle.render(_frame(cam, 640, 360), want_aux=True)
    assert_hits_equal(aux_g, aux_c, f"orbit {i}", ignore_iters=macro)
    assert out_g.tobytes() == out_c.tobytes()


@pytest.mark.parametrize("bounces", [1, 2, 3])
def test_render_bounces(hash_scene, hash_oracle, shading_inputs, bounces):
    """Secondary diffuse rays with the blue-noise table and the sky cube: radiance (f16) bytes equal
    — tolerance 0 under the canonical arithmetic (stated tolerance of the contract: <= 1 f16 ulp)."""
    from scenes import camera

    (bn, _), (desc, tex, _) = shading_inputs
    ctx = ctx_for(hash_scene)
    ctx.set_blue_noise(bn)
    ctx.set_sky(desc, tex)
    hash_oracle.set_blue_noise(bn)
    hash_oracle.set_sky(desc, tex)
    cam = camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45)
    for frame_no in (1, 2, 77):
        out_g, aux_g = ctx.render(_frame(cam, 320, 180, bounces=bounces, frame_no=frame_no), want_aux=True)
        out_c, aux_c, st = hash_oracle.render(_frame(cam, 320, 180, bounces=bounces, frame_no=frame_no), want_aux=True)
        assert st.rays > 320 * 180
        assert_hits_equal(aux_g, aux_c, f"bounces {bounces} primary aux", ignore_iters=True)
        for k in ("albedo", "depth", "irr_rg", "irr_bx"):
            a, b = out_g[k].view(np.uint32), out_c[k].view(np.uint32)
            assert np.array_equal(a, b), f"{k}: {np.count_nonzero(a != b)} texels differ (bounces {bounces}, frame {frame_no})"
    ctx.close()


@pytest.mark.parametrize("size", [(320, 180), (36, 4), (260, 148)])
def test_bounce_compaction_equals_per_pixel_path(bench_ctx, bench_oracle, shading_inputs, size):
    """Frames with bounces: the wavefront form (camera pass, then per level a persistent lane-refilling trace pass and a shade pass) gives the
    bytes of the one-thread-per-pixel kernel and of the oracle, incl. aux records, partitions and odd frame sizes."""
    from scenes import camera
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    bench_ctx.set_blue_noise(bn)
    bench_ctx.set_sky(desc, tex)
    bench_oracle.set_blue_noise(bn)
    bench_oracle.set_sky(desc, tex)
    w, h = size
    cam = camera.orbit_cameras(4, seed=2)[1]
    try:
        for bounces in (1, 3):
            want, aux_c, _ = bench_oracle.render(_frame(cam, w, h, bounces=bounces, frame_no=7), want_aux=True)
            for mode in ("per-pixel", "wavefront"):
                bench_ctx.set_option("wavefront", int(mode == "wavefront"))
                got, aux_g = bench_ctx.render(_frame(cam, w, h, bounces=bounces, frame_no=7), want_aux=True)
                assert got.tobytes() == want.tobytes(), (size, bounces, mode)
                assert_hits_equal(aux_g, aux_c, f"bounce mode {mode}", ignore_iters=True)
                part = np.zeros_like(want)
                for p in range(3):
                    f = _frame(cam, w, h, bounces=bounces, frame_no=7, part_index=p, part_count=3, flags=capi.VRT_FRAME_PART_ROWS)
                    bench_ctx._chk(bench_ctx.lib.vrt_render(bench_ctx.h, __import__("ctypes").byref(f), part.ctypes.data, None))
                assert part.tobytes() == want.tobytes(), (size, bounces, mode, "band split")
    finally:
        bench_ctx.set_option("wavefront", 2)
    with pytest.raises(capi.VrtError):  # round 1's CTA-level compaction is gone
        bench_ctx.set_option("compact_bounces", 1)


def test_wavefront_self_tuning_keeps_frames_identical(bench_ctx, bench_oracle, shading_inputs):
    """Default setting: the first bounce frames after a scene change are timed in both forms and one is kept — every frame of
    the sequence, whichever form traced it, equals the oracle's, also across a sync that changes sector emptiness."""
    import torch

    from scenes import camera

    (bn, _), (desc, tex, _) = shading_inputs
    for m in (bench_ctx, bench_oracle):
        m.set_blue_noise(bn)
        m.set_sky(desc, tex)
    bench_ctx.set_option("wavefront", 2)
    cam = camera.orbit_cameras(4, seed=3)[2]
    w, h = 640, 360
    want, _, _ = bench_oracle.render(_frame(cam, w, h, bounces=2, frame_no=4), want_aux=False)
    fb = torch.zeros(w * h * 4, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    for k in range(6):
        bench_ctx.render_device(_frame(cam, w, h, bounces=2, frame_no=4), fb.data_ptr(), None, stream.cuda_stream)
        stream.synchronize()
        assert fb.cpu().numpy().tobytes() == want.tobytes(), f"frame {k}"
    # a new resident sector (emptiness changes) restarts the tuning
    brick = np.zeros((1, 512), np.uint8)
    brick[0, 7] = 250
    for m in (bench_ctx, bench_oracle):
        m.sync([(30, 9, 30, 1, 1, brick)])
    want2, _, _ = bench_oracle.render(_frame(cam, w, h, bounces=2, frame_no=4), want_aux=False)
    for k in range(4):
        bench_ctx.render_device(_frame(cam, w, h, bounces=2, frame_no=4), fb.data_ptr(), None, stream.cuda_stream)
        stream.synchronize()
        assert fb.cpu().numpy().tobytes() == want2.tobytes(), f"frame {k} after the edit"
    for m in (bench_ctx, bench_oracle):  # put the shared scene back
        m.sync([(30, 9, 30, 0, 1, None, True)])


def test_render_needs_blue_noise_for_bounces(hash_scene):
    from scenes import camera
    from voxelrt_b200 import capi

    ctx = ctx_for(hash_scene)
    with pytest.raises(capi.VrtError):
        ctx.render(_frame(camera.Camera(), 64, 64, bounces=1))
    ctx.close()


@pytest.mark.parametrize("parts", [2, 4, 8])
def test_render_tile_partition_union(hash_ctx, parts):
    """Screen-tile split: the union of the N partial frames is the 1-GPU frame, byte for
    byte, and the parts are disjoint."""
    from scenes import camera

    cam = camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45)
    w, h = 416, 260
    full, _ = hash_ctx.render(_frame(cam, w, h))
    full = full.view(np.uint32).reshape(-1, 64)
    acc = np.zeros_like(full)
    for p in range(parts):
        part, _ = hash_ctx.render(_frame(cam, w, h, part_index=p, part_count=parts))
        part = part.view(np.uint32).reshape(-1, 64)
        assert not np.any((acc != 0) & (part != 0))
        acc |= part
    assert np.array_equal(acc, full)


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_render_gather_row_bands(hash_ctx, parts):
    """vrt_render_gather (the pipelined multi-GPU exchange), exercised on one GPU: every "rank" renders its 8-pixel bands
    into its own buffer and the strided band copy lands them in the owner's framebuffer; after all ranks the owner holds
    the single-GPU frame byte for byte, for frame heights with and without a partial last band, with the local buffers
    alternating as a pipelined caller would."""
    import torch

    from scenes import camera
    from voxelrt_b200 import capi

    cam = camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45)
    stream = torch.cuda.Stream()
    for w, h in ((416, 260), (256, 128), (64, 36)):
        want, _ = hash_ctx.render(_frame(cam, w, h))
        owner = torch.zeros(w * h * 4, dtype=torch.int32, device="cuda")
        local = [torch.full((w * h * 4,), -1, dtype=torch.int32, device="cuda") for _ in range(capi.VRT_GATHER_DEPTH)]
        torch.cuda.synchronize()
        for p in range(parts):
            f = _frame(cam, w, h, part_index=p, part_count=parts, flags=capi.VRT_FRAME_PART_ROWS)
            hash_ctx.render_gather(f, local[p % capi.VRT_GATHER_DEPTH].data_ptr(), owner.data_ptr(), stream.cuda_stream)
        hash_ctx.gather_wait(stream.cuda_stream)
        stream.synchronize()
        assert owner.cpu().numpy().tobytes() == want.tobytes(), (w, h, parts)
        # host-buffer form of the band split: every rank's call fills exactly its bands of the caller's frame
        host = np.zeros_like(want)
        for p in range(parts):
            f = _frame(cam, w, h, part_index=p, part_count=parts, flags=capi.VRT_FRAME_PART_ROWS)
            hash_ctx._chk(hash_ctx.lib.vrt_render(hash_ctx.h, __import__("ctypes").byref(f), host.ctypes.data, None))
        assert host.tobytes() == want.tobytes(), (w, h, parts, "host bands")
    with pytest.raises(capi.VrtError):  # the tile split cannot be moved with band copies
        hash_ctx.render_gather(_frame(cam, 64, 36, part_index=0, part_count=2), local[0].data_ptr(), owner.data_ptr(), stream.cuda_stream)


@pytest.mark.parametrize("parts", [1, 3, 8])
def test_compact_output_is_the_varying_half_of_the_tile_framebuffer(hash_ctx, parts):
    """VRT_FRAME_COMPACT (primary-only frames, 8 B/px): albedo + depth equal those of the full 16 B/px frame, whose irradiance words are the
    constant 0x3C003C00 — through vrt_render (whole frame, band split into one host frame), the linear form, and vrt_render_gather."""
    import ctypes as C

    import torch

    from scenes import camera
    from voxelrt_b200 import capi

    cam = camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45)
    stream = torch.cuda.Stream()
    for w, h in ((1280, 720), (416, 260), (64, 36)):
        full, _ = hash_ctx.render(_frame(cam, w, h))
        assert (full["irr_rg"] == 0x3C003C00).all() and (full["irr_bx"] == 0x3C003C00).all()
        host = np.zeros(w * h // 16, capi.TILE_AD_DTYPE)
        for p in range(parts):
            f = _frame(cam, w, h, part_index=p, part_count=parts, flags=capi.VRT_FRAME_COMPACT | (capi.VRT_FRAME_PART_ROWS if parts > 1 else 0))
            hash_ctx._chk(hash_ctx.lib.vrt_render(hash_ctx.h, C.byref(f), host.ctypes.data, None))
        assert np.array_equal(host["albedo"], full["albedo"]) and host["depth"].tobytes() == full["depth"].tobytes(), (w, h, parts)
        owner = torch.zeros(w * h * 2, dtype=torch.int32, device="cuda")
        local = [torch.full((w * h * 2,), -1, dtype=torch.int32, device="cuda") for _ in range(capi.VRT_GATHER_DEPTH)]
        torch.cuda.synchronize()
        for p in range(parts):
            f = _frame(cam, w, h, part_index=p, part_count=parts, flags=capi.VRT_FRAME_COMPACT | capi.VRT_FRAME_PART_ROWS)
            hash_ctx.render_gather(f, local[p % capi.VRT_GATHER_DEPTH].data_ptr(), owner.data_ptr(), stream.cuda_stream)
        hash_ctx.gather_wait(stream.cuda_stream)
        stream.synchronize()
        assert owner.cpu().numpy().tobytes() == host.tobytes(), (w, h, parts, "gather")
    lin, _ = hash_ctx.render(_frame(cam, 132, 68, flags=capi.VRT_FRAME_LINEAR_OUTPUT | capi.VRT_FRAME_COMPACT))
    lin_full, _ = hash_ctx.render(_frame(cam, 132, 68, flags=capi.VRT_FRAME_LINEAR_OUTPUT))
    assert lin.shape == (2, 68, 132) and np.array_equal(lin, lin_full[:2])
    with pytest.raises(capi.VrtError):  # with bounces the irradiance is not constant
        hash_ctx.render(_frame(cam, 64, 36, bounces=1, flags=capi.VRT_FRAME_COMPACT))


# ---------------------------------------------------------------------------------------------
# hit query: vrt_hit_query == VoxelMap::RayCast (VoxelMap.cpp:140-170), fp64
# ---------------------------------------------------------------------------------------------
def test_hit_query(hash_ctx, hash_oracle):
    rng = np.random.default_rng(3)
    n = 50_000
    o = np.stack([rng.uniform(0, 192, n), rng.uniform(0, 128, n), rng.uniform(0, 192, n)], 1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    got = hash_ctx.hit_query(o, d)
    want = hash_oracle.hit_query(o, d)
    assert (want["dist"] >= 0).sum() > 5000
    for name in got.dtype.names:
        a, b = got[name], want[name]
        if a.dtype.kind == "f":
            a, b = a.view(f"u{a.dtype.itemsize}"), b.view(f"u{b.dtype.itemsize}")
        assert np.array_equal(a, b), name


# ---------------------------------------------------------------------------------------------
# golden vectors computed by the REFERENCE ITSELF (tests/golden, made from oracle/_ref)
# ---------------------------------------------------------------------------------------------
def test_gpu_trace_matches_reference_golden(hash_scene, hash_ctx, macro):
    """vrt_trace against RayCast of the reference's own CpuRenderer.cpp (lane-wise), every field."""
    from pathlib import Path

    from scenes import terrain

    z = np.load(Path(__file__).resolve().parent / "golden" / "ref_trace_lane.npz")
    assert str(z["scene_digest"]) == terrain.scene_digest(hash_scene)
    want = z["hits"]
    got = hash_ctx.trace(z["origin"], z["dir"], z["world_origin"])
    hit = (want["flags"] & 0x100) != 0
    assert np.array_equal(got["flags"] & 0x13F, want["flags"] & 0x13F)
    assert np.array_equal(got["material"], want["material"])
    for f in ("dist", "px", "py", "pz", "u", "v"):
        ok = (got[f].view(np.uint32) == want[f].view(np.uint32)) | (np.isnan(got[f]) & np.isnan(want[f]))
        assert ok.all(), f
    for f in ("vx", "vy", "vz"):
        assert np.array_equal(got[f][hit], want[f][hit]), f


def test_gpu_hit_query_matches_reference_golden(hash_scene, hash_ctx):
    """vrt_hit_query against the reference's VoxelMap::RayCast (fp64), bit-exact."""
    from pathlib import Path

    z = np.load(Path(__file__).resolve().parent / "golden" / "ref_hit_query.npz")
    want = z["hits"]
    got = hash_ctx.hit_query(z["origin"], z["dir"])
    hit = want["dist"] >= 0
    assert np.array_equal(got["dist"].view(np.uint64), want["dist"].view(np.uint64))
    for f in ("nx", "ny", "nz", "u", "v"):
        assert np.array_equal(got[f][hit].view(np.uint32), want[f][hit].view(np.uint32)), f
    for f in ("vx", "vy", "vz"):
        assert np.array_equal(got[f][hit], want[f][hit]), f
