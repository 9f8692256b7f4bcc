"""CPU tests of the image-space oracle (oracle/vrt_post_oracle.c: the reference's GBuffer shaders restated) — its
arithmetic helpers, the blit, and the properties the reference's denoiser has by construction.  The GLSL itself
cannot run here (no GL device), so these pin the restatement against independent numpy statements of the shader
formulas and against a committed regression fixture; the GPU parity tests then compare the CUDA path with it."""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import pytest

from scenes import gbuffer_synth as pu

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def pp():
    from oracle import pypostoracle

    return pypostoracle


def test_exp_log_f16_helpers(pp):
    lib = pp.load()
    xs = np.linspace(-86.9, 10, 4001).astype(np.float32)
    assert max(abs(lib.post_exp(float(x)) / math.exp(float(x)) - 1) for x in xs) < 3e-7
    assert lib.post_exp(0.0) == 1.0 and lib.post_exp(-87.0) == 0.0 and lib.post_exp(-1e30) == 0.0
    assert math.isnan(lib.post_exp(float("nan")))
    xs = np.exp(np.linspace(-80, 5, 4001)).astype(np.float32)
    assert max(abs(lib.post_log(float(x)) - math.log(float(x))) / max(1e-3, abs(math.log(float(x)))) for x in xs) < 3e-7
    assert lib.post_log(1.0) == 0.0
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.normal(size=4000) * 10 ** rng.uniform(-9, 4.5, 4000), [0, -0.0, 65504, 65519.99, 65520, 6e-8, 2.98e-8, 2.99e-8]])
    v = v.astype(np.float32)
    got = np.array([lib.post_f32_to_f16(float(x)) for x in v], np.uint16)
    with np.errstate(over="ignore"):
        assert np.array_equal(got, v.astype(np.float16).view(np.uint16))
    allh = np.arange(65536, dtype=np.uint16)
    f = np.array([lib.post_f16_to_f32(int(x)) for x in allh], np.float32)
    ref = allh.view(np.float16).astype(np.float32)
    assert not ((f.view(np.uint32) != ref.view(np.uint32)) & ~(np.isnan(f) & np.isnan(ref))).any()


SIZE = (64, 48)


def test_blit_matches_the_shader_formulas(pp):
    """CopyTiledFramebuffer.comp:10-37: tile addressing, sky pixels get white albedo, normal code moves to alpha."""
    w, h = SIZE
    seq = pu.synthetic_sequence(w, h, 1, seed=5)
    proj, inv, pos, tiles = seq[0]
    orc = pp.PostOracle(w, h)
    orc.set_camera(proj, inv, pos)
    orc.set_debug_channel(4)  # TraversalIters: blit + present only
    orc.denoise_present(tiles)
    a = pu.untile(tiles, w, h, "albedo")
    d = pu.untile(tiles, w, h, "depth")
    want_a = np.where(d < 0, (a & 0xFF000000) | 0xFFFFFF, a)
    assert np.array_equal(orc.read(orc.ALBEDO).reshape(h, w), want_a)
    assert np.array_equal(orc.read(orc.DEPTH).reshape(h, w), d)
    irr = orc.read(orc.IRR).reshape(h, w, 4)
    rg, bx = pu.untile(tiles, w, h, "irr_rg"), pu.untile(tiles, w, h, "irr_bx")
    assert np.array_equal(irr[..., 0], rg & 0xFFFF) and np.array_equal(irr[..., 1], rg >> 16)
    assert np.array_equal(irr[..., 2], bx & 0xFFFF) and not irr[..., 3].any()


def test_first_frame_has_no_history_and_static_camera_accumulates(pp):
    """Reproject.comp: history length 0 on the first frame, then +1 per frame up to 64 on a static view; sky stays 0."""
    w, h = SIZE
    seq = pu.synthetic_sequence(w, h, 70, seed=6, moving=False)
    orc = pp.PostOracle(w, h)
    orc.set_passes(0)
    sky = pu.untile(seq[0][3], w, h, "depth") < 0
    for f, (proj, inv, pos, tiles) in enumerate(seq):
        orc.set_camera(proj, inv, pos)
        orc.denoise_present(tiles)
        hist = orc.read(orc.HIST).reshape(h, w)
        assert not hist[sky].any()
        assert (hist[~sky] == min(f, 64)).all(), f"frame {f}: {np.unique(hist[~sky])}"


def test_temporal_accumulation_converges_to_the_mean(pp):
    """With a static camera the accumulated irradiance is the running mean of the noisy inputs (blend 1/(n+1))."""
    w, h = SIZE
    seq = pu.synthetic_sequence(w, h, 24, seed=7, moving=False)
    orc = pp.PostOracle(w, h)
    orc.set_passes(0)
    sky = pu.untile(seq[0][3], w, h, "depth") < 0
    acc = []
    for proj, inv, pos, tiles in seq:
        orc.set_camera(proj, inv, pos)
        orc.denoise_present(tiles)
        rg = pu.untile(tiles, w, h, "irr_rg")
        acc.append(pu.f16_to_f32(rg & 0xFFFF))
    mean = np.mean(acc[1:], axis=0)  # frame 0 has no history; frame 1 blends with factor 1/(0+1) and so replaces it
    got = pu.f16_to_f32(orc.read(orc.PREV_IRR).reshape(h, w, 4)[..., 0])  # N = 0: history = the reprojected plane
    inner = ~sky
    inner[:, 0] = False  # ivec2() truncates toward zero: a history position of -1e-5 becomes texel 0 with fract ~1 (Reproject.comp:51-52)
    assert np.abs(got - mean)[inner].max() < 0.01 * max(1.0, mean.max())
    # and the filtered path reduces the frame-to-frame noise further
    orc5 = pp.PostOracle(w, h)
    for proj, inv, pos, tiles in seq[:6]:
        orc5.set_camera(proj, inv, pos)
        orc5.denoise_present(tiles)
    filt = pu.f16_to_f32(orc5.read(orc5.IRR).reshape(h, w, 4)[..., 0])
    raw = acc[5]
    body = ~sky
    body[:, :2] = body[:, -2:] = False
    dx_f = np.abs(np.diff(filt, axis=1))[body[:, 1:]].mean()
    dx_r = np.abs(np.diff(raw, axis=1))[body[:, 1:]].mean()
    assert dx_f < 0.35 * dx_r, (dx_f, dx_r)


def test_constant_irradiance_is_a_fixed_point(pp):
    """Every filter stage is a normalised weighted mean: a constant irradiance comes out unchanged, variance 0."""
    w, h = SIZE
    seq = pu.synthetic_sequence(w, h, 4, seed=8, moving=True)
    orc = pp.PostOracle(w, h)
    for proj, inv, pos, tiles in seq:
        t = tiles.copy()
        t["irr_rg"][:] = 0x38003C00  # (1.0, 0.5)
        t["irr_bx"][:] = 0x3400  # 0.25
        orc.set_camera(proj, inv, pos)
        orc.denoise_present(t)
        for plane in (orc.IRR, orc.PREV_IRR):
            irr = orc.read(plane).reshape(h, w, 4)
            assert (irr[..., 0] == 0x3C00).all() and (irr[..., 1] == 0x3800).all() and (irr[..., 2] == 0x3400).all()
            assert (pu.f16_to_f32(irr[..., 3]) < 1e-3).all()


def test_reset_history_caps_the_length_and_normal_change_drops_it(pp):
    w, h = SIZE
    seq = pu.synthetic_sequence(w, h, 12, seed=9, moving=False)
    orc = pp.PostOracle(w, h)
    orc.set_passes(1)
    sky = pu.untile(seq[0][3], w, h, "depth") < 0
    for f, (proj, inv, pos, tiles) in enumerate(seq[:10]):
        orc.set_camera(proj, inv, pos)
        orc.denoise_present(tiles)
    assert (orc.read(orc.HIST).reshape(h, w)[~sky] == 9).all()
    proj, inv, pos, tiles = seq[10]
    orc.set_camera(proj, inv, pos, reset_history=True)  # u_ForceResetHistory: min(len, 6) + 1
    orc.denoise_present(tiles)
    assert (orc.read(orc.HIST).reshape(h, w)[~sky] == 7).all()
    # flip the normals of a block: dot(n, n') < 0.5 there -> every history sample is rejected
    t = seq[11][3].copy()
    tv = t.view(pu.TILE_DTYPE).reshape(h // 4, w // 4)
    tv["albedo"][6:9, 5:9] = (tv["albedo"][6:9, 5:9] & 0x00FFFFFF) | (pu.normal_code(1, 0, 0) << 24)
    orc.set_camera(*seq[11][:3])
    orc.denoise_present(t)
    hist = orc.read(orc.HIST).reshape(h, w)
    assert not hist[24:36, 20:36].any() and (hist[40:, :16] == 8).all()


@pytest.mark.parametrize("passes", [0, 1, 2, 3, 4, 5])
def test_pass_rotation_follows_gbuffer_h(pp, passes):
    """GBuffer.h:100-130: which buffer carries which name after a frame.  N = 0: Prev = the reprojected plane;
    N >= 1: Prev = output of à-trous pass 0; N = 2 presents the PREVIOUS frame's history (the FIXME at :121)."""
    w, h = SIZE
    seq = pu.synthetic_sequence(w, h, 3, seed=10 + passes)
    orc = pp.PostOracle(w, h)
    orc.set_passes(passes)
    orc.set_debug_channel(2)
    prev_hist = None
    for proj, inv, pos, tiles in seq:
        orc.set_camera(proj, inv, pos)
        img = orc.denoise_present(tiles)
        shown = orc.read(orc.IRR)
        assert img.shape == (h, w) and (img >> 24 == 255).all()
        if passes == 2 and prev_hist is not None:
            assert np.array_equal(shown, prev_hist)
        prev_hist = orc.read(orc.PREV_IRR)
        sky = (pu.untile(tiles, w, h, "depth") < 0).reshape(-1)
        # sky pixels pass through every stage untouched (Filter.comp:21-24,74-77): history == this frame's input there
        rg = pu.untile(tiles, w, h, "irr_rg").reshape(-1)
        assert np.array_equal(prev_hist[sky, 0], (rg & 0xFFFF)[sky].astype(np.uint16))


def test_present_matches_the_formula_in_float64(pp):
    """GBufferBlit.frag:18-30: albedo * irradiance * 0.48 -> ACES -> pow 0.45, within one 8-bit step of float64."""
    w, h = SIZE
    seq = pu.synthetic_sequence(w, h, 1, seed=20)
    proj, inv, pos, tiles = seq[0]
    orc = pp.PostOracle(w, h)
    orc.set_passes(0)
    orc.set_camera(proj, inv, pos)
    img = orc.denoise_present(tiles)
    a = orc.read(orc.ALBEDO).reshape(h, w)
    irr = pu.f16_to_f32(orc.read(orc.PREV_IRR).reshape(h, w, 4)).astype(np.float64)
    for ch in range(3):
        alb = ((a >> (8 * ch)) & 255) / 255.0
        v = alb * irr[..., ch] * 0.48 * 0.6
        c = np.clip((v * (2.51 * v + 0.03)) / (v * (2.43 * v + 0.59) + 0.14), 0, 1) ** 0.45
        got = ((img >> (8 * ch)) & 255).astype(np.int64)
        assert np.abs(got - np.floor(c * 255 + 0.5)).max() <= 1


def test_debug_channels(pp):
    w, h = SIZE
    proj, inv, pos, tiles = pu.synthetic_sequence(w, h, 1, seed=21)[0]
    a = pu.untile(tiles, w, h, "albedo")
    d = pu.untile(tiles, w, h, "depth")
    for ch in (1, 3):
        orc = pp.PostOracle(w, h)
        orc.set_debug_channel(ch)
        orc.set_camera(proj, inv, pos)
        img = orc.denoise_present(tiles)
        if ch == 1:
            assert np.array_equal(img & 0xFFFFFF, np.where(d < 0, 0xFFFFFF, a & 0xFFFFFF))
        else:
            code = a >> 24
            for k in range(3):
                n = ((code >> (2 * k)) & 3).astype(np.int64) - 1
                assert np.array_equal((img >> (8 * k)) & 255, np.floor((n * 0.5 + 0.5) * 255 + 0.5).astype(np.uint32))


def test_regression_fixture(pp):
    """The committed fixture (tests/golden/post_sequence.npz, written by tests/golden/make_post_fixture.py from this
    oracle) detects any change of the canonical arithmetic; the GPU tests compare the CUDA path with the same file."""
    z = np.load(GOLDEN / "post_sequence.npz")
    w, h, frames, passes = (int(z[k]) for k in ("w", "h", "frames", "passes"))
    seq = pu.synthetic_sequence(w, h, frames, seed=int(z["seed"]))
    orc = pp.PostOracle(w, h)
    orc.set_passes(passes)
    for f, (proj, inv, pos, tiles) in enumerate(seq):
        assert np.array_equal(tiles.view(np.uint32), z["tiles"][f]), "the synthetic sequence generator changed"
        orc.set_camera(proj, inv, pos, reset_history=(f == 3))
        img = orc.denoise_present(tiles)
        assert np.array_equal(img, z["rgba"][f]), f"frame {f}"
    assert np.array_equal(orc.read(orc.HIST), z["hist"]) and np.array_equal(orc.read(orc.MOMENTS), z["moments"])
    assert np.array_equal(orc.read(orc.PREV_IRR), z["prev_irr"])


@pytest.mark.parametrize("passes,moving", [(5, True), (0, True), (2, True), (3, False)])
def test_oracle_agrees_with_an_independent_float64_model(pp, passes, moving):
    """tests/gbuffer_model_np.py restates the same shaders a second time (numpy, float64, vectorised, written from the GLSL).
    Threshold tests (plane distance, bounds after truncation, wsum) can flip on rounding, so a small fraction of pixels may
    differ; everywhere else the planes agree to f16 precision and the image to one 8-bit step."""
    from gbuffer_model_np import NumpyGBuffer

    w, h = SIZE
    seq = pu.synthetic_sequence(w, h, 5, seed=40 + passes, moving=moving)
    orc, mdl = pp.PostOracle(w, h), NumpyGBuffer(w, h)
    orc.set_passes(passes)
    for f, (proj, inv, pos, tiles) in enumerate(seq):
        reset = f == 3
        orc.set_camera(proj, inv, pos, reset_history=reset)
        mdl.set_camera(proj, inv, pos)
        img = orc.denoise_present(tiles)
        rgb = mdl.frame(tiles, reset=reset, passes=passes)
        hist_o = orc.read(orc.HIST).reshape(h, w).astype(np.int64)
        same_hist = hist_o == mdl.hist
        assert same_hist.mean() > 0.999, (f, same_hist.mean())
        for plane_o, plane_m, name in ((orc.IRR, mdl.irr, "irr"), (orc.PREV_IRR, mdl.prev_irr, "prev"), (orc.TEMP_IRR, mdl.temp_irr, "temp")):
            got = pu.f16_to_f32(orc.read(plane_o).reshape(h, w, 4)).astype(np.float64)
            err = np.abs(got - plane_m) / np.maximum(1e-2, np.abs(plane_m))
            close = (err < 4e-3).all(axis=-1)
            # (static camera: every reprojected position sits exactly on a pixel centre, so float32-vs-float64 rounding decides the
            # bilinear taps along the left / top edge columns — a few more threshold flips than with a moving camera)
            assert close.mean() > (0.985 if moving else 0.95), (f, name, close.mean(), float(np.nanmax(err)))
        mom = pu.f16_to_f32(orc.read(orc.MOMENTS).reshape(h, w, 2)).astype(np.float64)
        assert (np.abs(mom - mdl.moments) < 4e-3 * np.maximum(1.0, np.abs(mdl.moments))).all(axis=-1).mean() > 0.985
        ch = np.stack([(img >> s) & 255 for s in (0, 8, 16)], -1).astype(np.int64)
        assert (np.abs(ch - rgb) <= 1).all(axis=-1).mean() > 0.999, f
