"""WHOLE FRAMES of the CUDA path against the REFERENCE: the golden G-buffers the reference's own RenderRow rendered (tests/golden/ref_frames.npz)
and — where oracle/_ref can run (an AVX-512 host, as the GPU boxes are) — frames rendered live by it.  Byte equality of every plane: albedo +
normal, depth, irradiance, through the C ABI (vrt_render), in both forms of the bounce path (one thread per pixel / wavefront)."""
from __future__ import annotations

import numpy as np
import pytest

import golden_frames as gf
from conftest import ctx_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    z = gf.load()
    if not gf.assets_match(z):
        pytest.skip("scenes/_ref does not hold the blue-noise table / sky cube the golden frames were rendered with")
    return z


def _shaded_ctx(scene, shading_inputs, **kw):
    (bn, _), (desc, tex, _) = shading_inputs
    ctx = ctx_for(scene, **kw)
    ctx.set_blue_noise(bn)
    ctx.set_sky(desc, tex)
    return ctx


def test_gpu_reproduces_reference_rendered_frames(golden, hash_scene, shading_inputs):
    from scenes import terrain

    assert terrain.scene_digest(hash_scene) == str(golden["hash_scene_digest"])
    ctx = _shaded_ctx(hash_scene, shading_inputs)
    for wave in (0, 1):
        ctx.set_option("wavefront", wave)
        for name in gf.HASH_NAMES:
            got, _ = ctx.render(gf.frame_of(golden, name))
            gf.assert_tiles_equal(got, golden[name + "_tiles"], f"{name} (wavefront={wave})")
    ctx.close()


def test_gpu_reproduces_reference_bench_frames(golden, bench_scene, shading_inputs):
    """BASELINE configs[0] / [1]: 1280x720 (0 and 1 bounce) and the full 3840x2160 primary frame — 8,294,400 pixels — equal the frames the
    reference rendered (SHA-256 per plane)."""
    from scenes import terrain

    if terrain.scene_digest(bench_scene) != str(golden["bench_scene_digest"]):
        pytest.skip("bench terrain differs from the one the golden frames were rendered on")
    ctx = _shaded_ctx(bench_scene, shading_inputs)
    for wave in (0, 1):
        ctx.set_option("wavefront", wave)
        for name in gf.BENCH_NAMES:
            got, _ = ctx.render(gf.frame_of(golden, name))
            gf.assert_digests_equal(got, golden, name)
    ctx.close()


def test_gpu_frames_equal_live_reference(bench_scene, shading_inputs):
    """Live: the reference's RenderRow (oracle/_ref, all host threads) against the CUDA path on frames no fixture holds — 1920x1080 with 2
    bounces from two cameras."""
    from oracle import refharness
    from scenes import camera, terrain
    from voxelrt_b200 import capi

    if not refharness.available():
        pytest.skip("oracle/_ref/libref_cpu.so cannot run here")
    (bn, _), (desc, tex, _) = shading_inputs
    ref = refharness.RefMap()
    ref.set_palette(bench_scene["palette"])
    ref.sync(terrain.scene_records(bench_scene))
    ref.set_blue_noise(bn)
    ref.set_sky(desc, tex)
    ctx = _shaded_ctx(bench_scene, shading_inputs)
    w, h = 1920, 1080
    for k, cam in enumerate((camera.Camera(), camera.orbit_cameras(4, seed=5)[1])):
        proj, inv, wo, frac = cam.matrices(w, h)
        want, _ = ref.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=11 + k, bounces=2))
        for wave in (0, 1):
            ctx.set_option("wavefront", wave)
            got, _ = ctx.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=11 + k, bounces=2))
            gf.assert_tiles_equal(got, want, f"1080p 2 bounces camera {k} wavefront={wave}")
    ref.close()
    ctx.close()


def test_gpu_sponza_frame_equals_live_reference(shading_inputs):
    """BASELINE configs[2]'s scene in the size the reference's CPU view holds (Sponza voxelised into 1024^3): one bounce, 1280x720, inside the
    atrium — against the reference's own RenderRow."""
    from oracle import refharness
    from scenes import camera, models, terrain
    from voxelrt_b200 import capi

    if not refharness.available() or not models.sponza_available(1024):
        pytest.skip("needs oracle/_ref and scenes/_ref/sponza_1024.dat")
    scene = models.sponza(1024)
    (bn, _), (desc, tex, _) = shading_inputs
    ref = refharness.RefMap()
    ref.set_palette(scene["palette"])
    ref.sync(terrain.scene_records(scene))
    ref.set_blue_noise(bn)
    ref.set_sky(desc, tex)
    ctx = _shaded_ctx(scene, shading_inputs)
    w, h = 1280, 720
    proj, inv, wo, frac = camera.Camera(pos=(210.3, 80.2, 505.7), yaw=1.5, pitch=-0.15).matrices(w, h)
    want, _ = ref.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=2, bounces=1))
    for wave in (0, 1):
        ctx.set_option("wavefront", wave)
        got, _ = ctx.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=2, bounces=1))
        gf.assert_tiles_equal(got, want, f"sponza 1024 wavefront={wave}")
    ref.close()
    ctx.close()
