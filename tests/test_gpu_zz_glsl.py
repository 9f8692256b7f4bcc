"""GPU parity of vrt_trace_glsl (the GLSL renderer's rayCast / rayCastCoarse semantics, include/voxelrt_b200.h) against the
oracle's orc_trace_glsl: every VrtHit field bit-exact (both sides spell the same fp32 operations, no FMA)."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_hits_equal, ctx_for, random_rays
from test_glsl_oracle import camera_frame_rays

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("flags", [0, 1, 2, 3])
def test_trace_glsl_matches_oracle(hash_scene, hash_oracle, flags):
    ctx = ctx_for(hash_scene)
    for k in range(4):
        wo, o, d = camera_frame_rays(20000, 900 + k)
        want, _ = hash_oracle.trace_glsl(o, d, wo, flags)
        got = ctx.trace_glsl(o, d, wo, flags)
        assert_hits_equal(got, want, f"flags={flags} wo={wo}")
    # origins far from the world origin (bounce-ray like) and outside the view box
    wo = (96, 64, 96)
    o, d = random_rays(np.random.default_rng(7), 30000, 192, 128, wo)
    want, _ = hash_oracle.trace_glsl(o, d, wo, flags)
    assert_hits_equal(ctx.trace_glsl(o, d, wo, flags), want, f"flags={flags} far origins")
    wo = (96, 600, 96)
    rng = np.random.default_rng(8)
    o = (rng.random((5000, 3)) - 0.5).astype(np.float32)
    d = rng.normal(size=(5000, 3)) * 0.08
    d[:, 1] = -1.0
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    want, _ = hash_oracle.trace_glsl(o, d, wo, flags)
    assert ((want["flags"] & 0x100) != 0).mean() > 0.5
    assert_hits_equal(ctx.trace_glsl(o, d, wo, flags), want, f"flags={flags} outside origins")
    ctx.close()


def test_trace_glsl_special_directions(hash_scene, hash_oracle):
    """Zero / negative-zero / denormal / huge / NaN direction components and NaN origins."""
    ctx = ctx_for(hash_scene)
    wo = (90, 70, 90)
    vals = np.array([0.0, -0.0, 1.0, -1.0, 1e-39, -1e-39, 1e30, np.inf, -np.inf, np.nan, 0.3, -0.7], np.float32)
    d = np.array([(a, b, c) for a in vals for b in vals for c in vals], np.float32)
    o = np.tile(np.array([[0.25, 0.5, 0.75]], np.float32), (len(d), 1))
    o[::7, 0] = np.nan
    for flags in (0, 1, 2):
        want, _ = hash_oracle.trace_glsl(o, d, wo, flags)
        assert_hits_equal(ctx.trace_glsl(o, d, wo, flags), want, f"special flags={flags}")
    ctx.close()


def test_trace_glsl_after_edits(hash_scene, hash_oracle):
    """The 128^3-level masks are rebuilt when a sync changes which sectors hold bricks."""
    from oracle import pyoracle
    from scenes import terrain
    from voxelrt_b200 import capi

    ctx = ctx_for(hash_scene)
    orc = pyoracle.OracleMap(6, 4)
    orc.set_palette(hash_scene["palette"])
    orc.sync(terrain.scene_records(hash_scene))
    wo, o, d = camera_frame_rays(20000, 950)
    assert_hits_equal(ctx.trace_glsl(o, d, wo, 0), orc.trace_glsl(o, d, wo, 0)[0], "before")
    # add a lone brick in an empty far sector, remove one populated sector
    brick = np.full((1, 512), 250, np.uint8)
    recs = [(40, 10, 40, 1 << 21, 1 << 21, brick), (2, 1, 2, 0, ~0 & 0xFFFFFFFFFFFFFFFF, None, True)]
    ctx.sync(recs)
    orc.sync(recs)
    for flags in (0, 1):
        assert_hits_equal(ctx.trace_glsl(o, d, wo, flags), orc.trace_glsl(o, d, wo, flags)[0], f"after edit flags={flags}")
    wo2 = (40 * 32 + 5, 10 * 32 + 60, 40 * 32 + 7)
    o2 = np.random.default_rng(2).random((4000, 3)).astype(np.float32)
    d2 = np.tile(np.array([[0.05, -1.0, 0.08]], np.float32), (4000, 1)) + np.random.default_rng(3).normal(size=(4000, 3)).astype(np.float32) * 0.1
    want, _ = orc.trace_glsl(o2, d2, wo2, 0)
    assert ((want["flags"] & 0x100) != 0).any()
    assert_hits_equal(ctx.trace_glsl(o2, d2, wo2, 0), want, "lone brick")
    with pytest.raises(capi.VrtError):
        ctx.trace_glsl(o2, d2, wo2, 8)
    ctx.close()
