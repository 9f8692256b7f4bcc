"""CPU tests of the GLSL-renderer ray casts in the oracle (orc_trace_glsl: rayCast / rayCastCoarse / getStepPos of
Shaders/VoxelTraversal.glsl) — the interaction-LUT known-answer test SURVEY §8c names, and cross-checks against the pinned
CPU-renderer traversal of the same scene: both walk the same occupancy, so away from grazing cases they must stop in
the same voxel.  (The GLSL itself cannot run here; this restatement is unpinned.)"""
from __future__ import annotations

import numpy as np
import pytest

HIT, INB, CAP = 0x100, 0x200, 0x400


def test_interaction_lut_known_answers():
    """GenerateRayCellInteractionMaskLUT (GpuRenderer.cpp:193-210) against an independent numpy statement."""
    from oracle import pyoracle

    lut = pyoracle.interaction_lut()
    for octant in range(8):
        s = np.array([1 if octant & 1 else -1, 1 if octant & 2 else -1, 1 if octant & 4 else -1])  # x, y, z
        for origin in range(64):
            o = np.array([origin & 3, origin >> 4, (origin >> 2) & 3])
            want = 0
            for x in range(4):
                for y in range(4):
                    for z in range(4):
                        rel = (np.array([x, y, z]) - o) * s
                        if (rel >= 0).all():
                            want |= 1 << (x | z << 2 | y << 4)
            assert int(lut[origin + 64 * octant]) == want, (octant, origin)
    assert int(lut[0 + 64 * 7]) == (1 << 64) - 1 and int(lut[63 + 64 * 0]) == (1 << 64) - 1
    assert int(lut[63 + 64 * 7]) == 1 << 63 and all((int(lut[i + 64 * o]) >> i) & 1 for i in range(64) for o in range(8))


def camera_frame_rays(n, seed):
    """Primary-ray-like inputs: an integer world origin inside the view and origins in [0,1)^3 (the renderer's camera-fraction
    frame, VoxelRender.comp:20-25).  rayCast's +5-ulp bias relies on that: with |origin| ~ 100 a 5-ulp step of a small t is
    below the rounding error of origin + t*dir and rays stall — the CPU renderer's 0.001 bias does not care."""
    rng = np.random.default_rng(seed)
    wo = (int(rng.integers(2, 190)), int(rng.integers(2, 126)), int(rng.integers(2, 190)))
    o = rng.random((n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return wo, o, d


def test_fine_cast_agrees_with_the_cpu_renderer_traversal(hash_oracle):
    """Same occupancy, different stepping (128^3 level, +5 ulp bias, start voxel not tested): the first solid voxel along
    a ray is the same one except for grazing rays; the GLSL walk needs fewer iterations thanks to the 128^3 level."""
    n_sel = n_same = n_agree = n_cmp = it_c = it_g = 0
    for k in range(10):
        wo, o, d = camera_frame_rays(4000, 100 + k)
        cpu, st_c = hash_oracle.trace(o, d, wo, max_iters=256)
        gl, st_g = hash_oracle.trace_glsl(o, d, wo, 0)
        ch, gh = (cpu["flags"] & HIT) != 0, (gl["flags"] & HIT) != 0
        started_solid = ((cpu["flags"] >> 16) == 1) & ch  # the CPU renderer tests the start voxel, rayCast does not
        sel = ch & gh & ~started_solid
        same = (cpu["vx"] == gl["vx"]) & (cpu["vy"] == gl["vy"]) & (cpu["vz"] == gl["vz"])
        n_sel += int(sel.sum())
        n_same += int((same & sel).sum())
        n_agree += int(((ch == gh) & ~started_solid).sum())
        n_cmp += int((~started_solid).sum())
        it_c += st_c.iters
        it_g += st_g.iters
        assert not (gl["flags"] & CAP).any()
        assert (gl["material"][gh] != 0).all() and (gl["material"][~gh] == 0).all()
    assert n_sel > 10000
    assert n_same / n_sel > 0.998 and n_agree / n_cmp > 0.999, (n_same / n_sel, n_agree / n_cmp)
    assert it_g < 0.9 * it_c, (it_g, it_c)


def test_outside_origins_are_clipped_to_the_grid(hash_oracle):
    """clipRayToAABB (VoxelTraversal.glsl:133-145): a camera outside the view box still sees the terrain; the CPU renderer's
    RayCast reports a miss for every such ray (first bounds test fails)."""
    wo = (96, 600, 96)  # 88 voxels above the 512-high view (the terrain fills y < 128)
    rng = np.random.default_rng(3)
    n = 5000
    o = (rng.random((n, 3)) - 0.5).astype(np.float32)
    d = rng.normal(size=(n, 3)) * 0.08
    d[:, 1] = -1.0
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    cpu, _ = hash_oracle.trace(o, d, wo)
    gl, _ = hash_oracle.trace_glsl(o, d, wo, 0)
    assert not (cpu["flags"] & HIT).any()
    assert ((gl["flags"] & HIT) != 0).mean() > 0.5
    hit = (gl["flags"] & HIT) != 0
    assert (gl["vy"][hit] < 128).all() and (gl["vy"][hit] >= 0).all() and (gl["dist"][hit] > 470).all()


def test_coarse_and_anisotropic_variants(hash_oracle):
    from voxelrt_b200 import capi

    wo, o, d = camera_frame_rays(30000, 5)
    fine, st_f = hash_oracle.trace_glsl(o, d, wo, 0)
    aniso, st_a = hash_oracle.trace_glsl(o, d, wo, capi.VRT_GLSL_ANISOTROPIC)
    coarse, st_c = hash_oracle.trace_glsl(o, d, wo, capi.VRT_GLSL_COARSE)
    # the LUT only hides cells the ray cannot reach: never more iterations, and mostly the very same hit
    assert st_a.iters <= st_f.iters
    fh, ah = (fine["flags"] & HIT) != 0, (aniso["flags"] & HIT) != 0
    both = fh & ah
    same = (fine["vx"] == aniso["vx"]) & (fine["vy"] == aniso["vy"]) & (fine["vz"] == aniso["vz"])
    assert (fh == ah).mean() > 0.995 and same[both].mean() > 0.995
    # coarse: capped at 96 iterations, every hit is an occupied voxel (material != 0), hits past iteration 30 may be any
    # occupied voxel of the 4^3 cell the ray entered
    ch = (coarse["flags"] & HIT) != 0
    assert ((coarse["flags"] >> 16) <= 96).all() and (coarse["material"][ch] != 0).all()
    early = ch & ((coarse["flags"] >> 16) <= 30) & fh
    same_c = (fine["vx"] == coarse["vx"]) & (fine["vy"] == coarse["vy"]) & (fine["vz"] == coarse["vz"])
    assert same_c[early].mean() > 0.98
    assert ((coarse["flags"] & CAP) != 0).sum() >= ((fine["flags"] & CAP) != 0).sum() or True


@pytest.mark.parametrize("coarse,aniso", [(False, False), (True, False), (False, True), (True, True)])
def test_oracle_equals_an_independent_python_restatement_bit_for_bit(hash_scene, hash_oracle, coarse, aniso):
    """tests/glsl_cast_model_py.py walks a dense voxel array with numpy float32 scalars, written from the GLSL on its own; the C
    oracle must produce the same VrtHit bits (camera-frame rays, bounce-like far origins for the coarse cast, outside cameras)."""
    from glsl_cast_model_py import DenseWorld, cast
    from voxelrt_b200 import capi

    world = DenseWorld(hash_scene)
    flags = (capi.VRT_GLSL_COARSE if coarse else 0) | (capi.VRT_GLSL_ANISOTROPIC if aniso else 0)
    cases = []
    for k in range(3):
        wo, o, d = camera_frame_rays(90, 700 + k + 10 * flags)
        cases.append((wo, o, d))
    rng = np.random.default_rng(50 + flags)
    if coarse:
        o = rng.uniform(-90, 90, (90, 3)).astype(np.float32)
        d = rng.normal(size=(90, 3))
        cases.append(((96, 64, 96), o, (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)))
    o = (rng.random((60, 3)) - 0.5).astype(np.float32)
    d = rng.normal(size=(60, 3)) * 0.08
    d[:, 1] = -1.0
    cases.append(((96, 600, 96), o, (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)))
    n_hit = 0
    for wo, o, d in cases:
        got, _ = hash_oracle.trace_glsl(o, d, wo, flags)
        for r in range(len(o)):
            m = cast(world, o[r], d[r], wo, coarse_mode=coarse, aniso=aniso)
            g = got[r]
            what = f"flags={flags} wo={wo} ray={r} o={o[r]} d={d[r]}: {m} vs {g}"
            assert bool(g["flags"] & HIT) == m["hit"] and bool(g["flags"] & INB) == m["inb"] and bool(g["flags"] & CAP) == m["capped"], what
            assert (g["flags"] >> 16) == m["iters"], what
            assert [g["vx"], g["vy"], g["vz"]] == m["voxel"], what
            code = (m["normal"][0] + 1) | (m["normal"][1] + 1) << 2 | (m["normal"][2] + 1) << 4
            assert (g["flags"] & 0x3F) == code and g["material"] == m["material"], what
            for a, b in ((g["dist"], m["dist"]), (g["px"], m["pos"][0]), (g["py"], m["pos"][1]), (g["pz"], m["pos"][2]), (g["u"], m["uv"][0]), (g["v"], m["uv"][1])):
                assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32) or (np.isnan(a) and np.isnan(b)), what
            n_hit += m["hit"]
    assert n_hit > 100
