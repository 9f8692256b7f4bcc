"""Scene ingest pinned against the reference's own code: scenes/models.py + scenes/voxelizer.c against
VoxelMap::VoxelizeModel (Voxelize.cpp:77-147) and glim::PaletteBuilder (Common/PaletteBuilder.h), both compiled from where they lie
into oracle/_ref/libref_cpu_strict.so (-ffp-contract=off on the unmodified sources: the voxeliser's edge tests are `>= 0` on fp32 sums,
so which products the compiler fuses decides single voxels; scenes/voxelizer.c fuses none, and neither does that build).

The model reaches both sides DECODED (triangles, texture coordinates, RGBA8 images): the reference's loaders sit on assimp and stb_image,
which are absent here.  Same palette, same voxels — byte for byte."""
from __future__ import annotations

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref():
    from oracle import refharness
    from scenes import models

    if not refharness.available("strict"):
        pytest.skip("oracle/_ref/libref_cpu_strict.so not available on this machine")
    if not models.VOX_LIB.exists():
        pytest.skip("scenes/_ref/libvoxelizer.so not built")
    return refharness


def _colors(rng, n, kind):
    if kind == "uniform":
        return rng.integers(0, 1 << 24, n).astype(np.uint32)
    base = rng.integers(0, 256, (12, 3))
    c = np.clip(base[rng.integers(0, 12, n)] + rng.normal(0, 14, (n, 3)), 0, 255).astype(np.uint32)
    return c[:, 0] | c[:, 1] << 8 | c[:, 2] << 16


@pytest.mark.parametrize("n,kind", [(1, "uniform"), (50, "uniform"), (300, "uniform"), (1200, "uniform"), (5000, "uniform"), (5000, "clustered"), (200_000, "clustered")])
def test_octree_palette_equals_palette_builder(ref, n, kind):
    """Same leaves after the cut to 240 (including the reference's over-deep cuts when a folded child was a subtree), same mean
    colours, same depth-first order, and the same nearest entry for every input colour."""
    from scenes import models

    c = _colors(np.random.default_rng(n), n, kind)
    rgb = np.stack([c & 255, (c >> 8) & 255, (c >> 16) & 255], axis=1).astype(np.uint8)
    want_pal, want_idx = ref.palette_build(c, 240, "strict")
    q = models.OctreePalette()
    q.add_colors(rgb[: n // 2])  # population and sums accumulate across calls
    q.add_colors(rgb[n // 2 :])
    got = q.build(240)
    assert got.shape[0] <= 240
    assert np.array_equal(got, want_pal)
    assert np.array_equal(models.nearest_palette_index(got, rgb), want_idx)


def _synthetic_model(seed, n_tris, span, tex_sizes=((16, 16), (64, 32), (4, 4), (128, 128))):
    rng = np.random.default_rng(seed)
    textures = []
    for w, h in tex_sizes:
        rgb = rng.integers(0, 1 << 24, (h, w)).astype(np.uint32)
        alpha = rng.choice([255, 255, 255, 210, 200, 199, 128, 127, 0], size=(h, w)).astype(np.uint32)  # both alpha thresholds (200, 128)
        textures.append(rgb | (alpha << 24))
    tris = rng.uniform(span[0], span[1], (n_tris, 3, 3)).astype(np.float32)
    uvs = rng.uniform(-1.5, 2.5, (n_tris, 3, 2)).astype(np.float32)  # Repeat addressing on both sides of [0, 1)
    tri_tex = rng.integers(0, len(textures), n_tris).astype(np.int32)
    # the cases the overlap test is touchy about: zero-area triangles (two and three equal corners, collinear corners: the reference does
    # not skip them, their NaN normal passes the slab test), axis-aligned triangles, corners on voxel corners
    tris[0, 1] = tris[0, 0]
    tris[1, 2] = tris[1, 1] = tris[1, 0]
    tris[2, :, 1] = 4.0
    tris[3, :, 0] = np.float32(2.5)
    tris[4] = np.round(tris[4])
    tris[5, 2] = (tris[5, 0] + tris[5, 1]) / 2
    return tris, uvs, tri_tex, textures


def _voxels(sectors):
    """{(sx, sy, sz, brick): 512 ids} without the all-empty bricks (the reference keeps bricks whose voxels were all written as id 0)."""
    out = {}
    for key, (mask, bricks) in sectors.items():
        j = 0
        for i in range(64):
            if mask >> i & 1:
                if bricks[j].any():
                    out[key + (i,)] = bricks[j]
                j += 1
    return out


def _assert_same_ingest(ref, tris, uvs, tri_tex, textures, size):
    from scenes import models

    scene, pal = models.voxelize_arrays(tris, uvs, tri_tex, textures, size, grid_bricks=size // 8 + 1)  # +1: corners ON the far faces
    rm = ref.RefMap("strict")
    try:
        rm.voxelize(tris, uvs, tri_tex, textures, size)
        mats = rm.materials()
        want = _voxels(rm.map_sectors())
    finally:
        rm.close()
    assert np.array_equal(np.array([m[:3] for m in mats[: len(pal)]], np.uint8), pal)
    assert all(m[:3] == (0, 0, 0) for m in mats[len(pal) :]), "palette length"
    got = _voxels(scene["sectors"])
    assert set(got) == set(want)
    assert all(np.array_equal(got[k], want[k]) for k in want)
    return sum(int((b != 0).sum()) for b in want.values())


@pytest.mark.parametrize("seed,n_tris,span,size", [(0, 60, (-3, 5), 64), (3, 200, (1, 9), 128), (4, 200, (-2, 2), 128)])
def test_synthetic_models_voxelise_like_the_reference(ref, seed, n_tris, span, size):
    """span (1, 9) lies wholly on the positive side: the reference measures such a model from ~0 (its running minimum starts at +FLT_MIN,
    Voxelize.cpp:101), and so do we."""
    solid = _assert_same_ingest(ref, *_synthetic_model(seed, n_tris, span), size)
    assert solid > 20_000


@pytest.mark.parametrize("size", [256, 1024])
def test_sponza_voxelises_like_the_reference(ref, size):
    """The app's start-up model (Main.cpp:42-47), all 262k triangles and 25 base-colour textures, decoded by OUR glTF / PIL readers."""
    from scenes import models

    if not models.REF_MODEL.exists():
        pytest.skip("reference assets absent")
    tris, uvs, mats, images = models.load_gltf(models.REF_MODEL)
    tex_list = sorted({p for p in images if p is not None})
    tex_index = {p: i for i, p in enumerate(tex_list)}
    textures = [models.load_rgba(p) for p in tex_list]
    tri_tex = np.array([tex_index[images[m]] for m in mats.tolist()], np.int32)
    solid = _assert_same_ingest(ref, tris, uvs, tri_tex, textures, size)
    assert solid > (400_000 if size == 256 else 8_000_000)
