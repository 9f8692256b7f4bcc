"""The CPU oracle against golden vectors produced by the REFERENCE ITSELF (tests/golden/*.npz, made by
tests/golden/make_golden.py from oracle/_ref = the reference's own CpuRenderer.cpp / VoxelMap.cpp).
Runs everywhere (no GPU, no /root/reference needed)."""
from __future__ import annotations

import hashlib
from pathlib import Path

import numpy as np
import pytest

from conftest import assert_hits_equal

GOLD = Path(__file__).resolve().parent / "golden"


def _load(name, hash_scene):
    from scenes import terrain

    z = np.load(GOLD / name)
    assert str(z["scene_digest"]) == terrain.scene_digest(hash_scene), "hash terrain changed: regenerate tests/golden"
    return z


def _assert_ref_fields(got, want, what, hit_only_voxel=True):
    """Fields the reference's VHitResult carries: material, distance, pos, normal, uv, hit mask."""
    hit_w = (want["flags"] & 0x100) != 0
    assert np.array_equal((got["flags"] & 0x100) != 0, hit_w), f"{what}: hit mask"
    assert np.array_equal(got["flags"] & 0x3F, want["flags"] & 0x3F), f"{what}: normal code"
    assert np.array_equal(got["material"], want["material"]), f"{what}: material"
    for f in ("dist", "px", "py", "pz", "u", "v"):
        a, b = got[f], want[f]
        ok = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        assert ok.all(), f"{what}: {f} differs at {np.count_nonzero(~ok)} rays"
    for f in ("vx", "vy", "vz"):
        assert np.array_equal(got[f][hit_w], want[f][hit_w]), f"{what}: {f}"


def test_oracle_trace_matches_reference_lanewise(hash_scene, hash_oracle):
    """RayCast (CpuRenderer.cpp:172-224), one active lane per packet: EVERY output bit equal."""
    z = _load("ref_trace_lane.npz", hash_scene)
    got, _ = hash_oracle.trace(z["origin"], z["dir"], z["world_origin"])
    _assert_ref_fields(got, z["hits"], "lane-wise")


def test_packet_coupled_quirks_are_the_only_difference(hash_scene, hash_oracle):
    """With 16 rays per packet the reference differs from its own lane-wise result only through the
    packet-coupled quirks: Q2 — a lane reaching the iteration cap corrupts the material
    of every lane of its packet; Q10 — a lane that stops in the FIRST iteration (ray starts in a solid
    voxel or outside the view) keeps being advanced by 0.001*dir while its packet goes on, which can
    move Pos/UV/voxel and, for inf/NaN/huge directions, even flip its hit mask.  Every other lane is
    bit-identical in every field."""
    z = _load("ref_trace_packet.npz", hash_scene)
    want = z["hits"]
    got, _ = hash_oracle.trace(z["origin"], z["dir"], z["world_origin"])
    iters = got["flags"] >> 16
    capped = (got["flags"] & 0x400) != 0
    packet_has_cap = np.repeat(capped.reshape(-1, 16).any(axis=1), 16)
    q10 = iters == 1
    clean = ~q10
    assert np.array_equal((got["flags"] & 0x13F)[clean], (want["flags"] & 0x13F)[clean]), "hit mask / normal outside Q10"
    for f in ("dist", "px", "py", "pz", "u", "v"):
        a, b = got[f], want[f]
        ok = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        assert ok[clean].all(), f"{f} differs outside Q10"
    mat_diff = got["material"] != want["material"]
    assert not (mat_diff & ~packet_has_cap & ~q10).any(), "material differs outside Q2/Q10"
    assert q10.sum() > 100 and packet_has_cap.sum() > 0  # the fixture does exercise both quirks


def test_oracle_hit_query_matches_reference(hash_scene, hash_oracle):
    """VoxelMap::RayCast (VoxelMap.cpp:140-170), fp64."""
    z = _load("ref_hit_query.npz", hash_scene)
    want = z["hits"]
    got = hash_oracle.hit_query(z["origin"], z["dir"])
    hit = want["dist"] >= 0
    assert hit.sum() > 500
    assert np.array_equal(got["dist"].view(np.uint64), want["dist"].view(np.uint64))
    for f in ("nx", "ny", "nz", "u", "v"):
        assert np.array_equal(got[f][hit].view(np.uint32), want[f][hit].view(np.uint32)), f
    for f in ("vx", "vy", "vz"):
        assert np.array_equal(got[f][hit], want[f][hit]), f


def test_oracle_storage_matches_reference(hash_scene, hash_oracle):
    """FlatVoxelStorage::SyncBuffers + UpdateOccupancy (CpuRenderer.cpp:33-83)."""
    z = _load("ref_storage.npz", hash_scene)
    for k, m, c in zip(z["keys"], z["masks"], z["cells"]):
        om, ob, oc = hash_oracle.read_sector(int(k[0]), int(k[1]), int(k[2]))
        assert om == int(m)
        assert np.array_equal(oc, c)


def test_material_and_pixel_formats_match_reference():
    """Material::GetEncoded (VoxelMap.h:27-41); RGBA8u / RG16f / R11G11B10f Pack (Texture.h)."""
    import ctypes as C

    from oracle import pyoracle

    z = np.load(GOLD / "ref_misc.npz")
    for m, e in zip(z["materials"], z["materials_encoded"]):
        assert pyoracle.encode_material(int(m[0]), int(m[1]), int(m[2]), int(m[3]), float(m[4])) == int(e)
    for c, e in zip(z["r11_in"], z["r11"]):
        assert pyoracle.pack_r11g11b10f(float(c[0]), float(c[1]), float(c[2])) == int(e)
    lib = pyoracle.load()
    lib.orc_pack_unorm8x4.argtypes = [C.c_float] * 4
    lib.orc_pack_unorm8x4.restype = C.c_uint32
    lib.orc_pack_half2.argtypes = [C.c_float] * 2
    lib.orc_pack_half2.restype = C.c_uint32
    for c, e in zip(z["rgba8_in"], z["rgba8"]):
        assert lib.orc_pack_unorm8x4(*[float(v) for v in c]) == int(e)
    for c, e in zip(z["rg16_in"], z["rg16"]):
        assert lib.orc_pack_half2(float(c[0]), float(c[1])) == int(e), (c, hex(int(e)))


def test_sincos_and_blue_noise_match_reference(hash_oracle):
    """simd::sincos_2pi (SIMD.h:175-190) bit-exact; VBlueNoise::Sample (CpuRenderer.cpp:254-270) exact."""
    import ctypes as C

    from oracle import pyoracle
    from scenes import shading

    z = np.load(GOLD / "ref_misc.npz")
    lib = pyoracle.load()
    lib.orc_sincos_2pi.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
    lib.orc_sincos_2pi.restype = None
    for x, (s, c) in zip(z["sincos_in"], z["sincos"]):
        a, b = C.c_float(), C.c_float()
        lib.orc_sincos_2pi(float(x), C.byref(a), C.byref(b))
        assert np.float32(a.value).view(np.uint32) == np.float32(s).view(np.uint32), x
        assert np.float32(b.value).view(np.uint32) == np.float32(c).view(np.uint32), x
    bn, _ = shading.load_blue_noise()
    if hashlib.sha256(bn.tobytes()).digest() != z["bn_table_sha"].tobytes():
        pytest.skip("blue-noise table differs from the one the golden was made with")
    hash_oracle.set_blue_noise(bn)
    out = (C.c_float * 2)()
    for q, tile in zip(z["bn_queries"], z["bn_tiles"]):
        x0, y0 = int(q[0]) & ~3, int(q[1]) & ~3
        for ty in range(4):
            for tx in range(4):
                lib.orc_blue_noise_sample(hash_oracle.h, x0 + tx, y0 + ty, int(q[2]), int(q[3]), out)
                assert np.float32(out[0]) == tile[ty, tx, 0] and np.float32(out[1]) == tile[ty, tx, 1]


# ---- whole frames the REFERENCE rendered (tests/golden/ref_frames.npz) ---------------------------------------------------------
def _golden_frames_or_skip():
    import golden_frames as gf

    z = gf.load()
    if not gf.assets_match(z):
        pytest.skip("scenes/_ref does not hold the blue-noise table / sky cube the golden frames were rendered with")
    return gf, z


@pytest.mark.parametrize("name", ["hash_b0", "hash_b1", "hash_b2", "hash_b3", "hash_inside_solid", "hash_outside_view"])
def test_oracle_reproduces_reference_rendered_frames(name, hash_scene, hash_oracle):
    """orc_render == the G-buffer the reference's own RenderRow produced (committed fixture), byte for byte: 0-3 bounces, a camera
    inside solid rock, a camera outside the view."""
    from scenes import shading, terrain

    gf, z = _golden_frames_or_skip()
    assert terrain.scene_digest(hash_scene) == str(z["hash_scene_digest"])
    hash_oracle.set_blue_noise(shading.load_blue_noise()[0])
    desc, texels, _ = shading.load_sky()
    hash_oracle.set_sky(desc, texels)
    got = hash_oracle.render(gf.frame_of(z, name))[0]
    gf.assert_tiles_equal(got, z[name + "_tiles"], name)


def test_oracle_reproduces_reference_bench_frames(bench_scene):
    """BASELINE configs[0] / [1]: the 1280x720 frame (0 and 1 bounce) and the 3840x2160 primary frame equal the reference-rendered
    ones (SHA-256 per plane; the frames are 15 / 133 MB)."""
    from oracle import pyoracle
    from scenes import shading, terrain

    gf, z = _golden_frames_or_skip()
    if terrain.scene_digest(bench_scene) != str(z["bench_scene_digest"]):
        pytest.skip("bench terrain differs from the one the golden frames were rendered on (no FastNoise2 here?)")
    orc = pyoracle.OracleMap(6, 4)
    orc.set_palette(bench_scene["palette"])
    orc.sync(terrain.scene_records(bench_scene))
    orc.set_blue_noise(shading.load_blue_noise()[0])
    desc, texels, _ = shading.load_sky()
    orc.set_sky(desc, texels)
    for name in gf.BENCH_NAMES:
        gf.assert_digests_equal(orc.render(gf.frame_of(z, name))[0], z, name)
    orc.close()
