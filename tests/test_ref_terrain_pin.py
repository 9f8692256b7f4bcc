"""The bench scenes' terrain pinned against the reference's own generator: src/VoxelRT/TerrainGenerator.cpp compiled from where it lies into
oracle/_ref/libref_terrain.so (with its vendored FastNoise2) and driven through its public interface — RequestSector / Poll, i.e. the worker
threads, GenerateSector and the copy of the non-empty bricks (TerrainGenerator.cpp:5-34,126-147).  scenes/terrain.py (numpy over the same
FastNoise2 library) and scenes/terrain_gen.c (the bulk generator of the 10 GB scene) must deliver the same allocation masks and the same voxels."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np
import pytest

LIB = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "libref_terrain.so"


@pytest.fixture(scope="module")
def ref_terrain():
    from oracle import refharness
    from scenes import terrain

    if not LIB.exists() or not refharness._cpu_ok():
        pytest.skip("oracle/_ref/libref_terrain.so not available on this machine")
    if not terrain.fastnoise_available():
        pytest.skip("scenes/_ref/libFastNoise.so not built")
    lib = C.CDLL(str(LIB))
    lib.ref_terrain_generate.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_terrain_generate.restype = C.c_int

    def generate(positions):
        pos = np.ascontiguousarray(positions, np.int32).reshape(-1, 3)
        masks = np.zeros(len(pos), np.uint64)
        bricks = np.zeros((len(pos), 64, 512), np.uint8)
        assert lib.ref_terrain_generate(len(pos), pos.ctypes.data, masks.ctypes.data, bricks.ctypes.data) == 0
        return masks, bricks

    return generate


def _assert_sector(mask, bricks, want_mask, want_bricks64, where):
    assert int(mask) == int(want_mask), where
    j = 0
    for b in range(64):
        if int(want_mask) >> b & 1:
            assert np.array_equal(bricks[j], want_bricks64[b]), (where, b)
            j += 1
    assert j == len(bricks)


def test_bench_terrain_is_the_reference_generators_output(ref_terrain):
    """BASELINE configs[1]'s scene, all 24 x 7 x 24 sectors (Main.cpp:63-69): every sector the reference generates is in our scene with
    the same mask and bricks, and every sector it finds empty is absent."""
    from scenes import terrain

    scene = terrain.terrain_fastnoise(24, 7, 24)
    pos = [(x, y, z) for y in range(7) for z in range(24) for x in range(24)]
    masks, bricks = ref_terrain(pos)
    n_bricks = 0
    for i, p in enumerate(pos):
        if masks[i] == 0:
            assert p not in scene["sectors"], p
            continue
        m, b = scene["sectors"][p]
        _assert_sector(m, b, masks[i], bricks[i], p)
        n_bricks += len(b)
    assert n_bricks == terrain.scene_stats(scene)["bricks"] > 100_000


def test_single_sectors_and_the_raised_terrain_of_the_large_scene(ref_terrain):
    """generate_sector_fastnoise on scattered sectors (far from the origin, negative coordinates), and the bulk generator of BASELINE
    configs[3] (scenes/terrain_gen.c, terrain raised by 512 voxels = 16 sectors): its sector (x, y, z) is the reference's (x, y - 16, z)."""
    from scenes import terrain

    rng = np.random.default_rng(4)
    pos = [(int(rng.integers(-40, 200)), int(rng.integers(-2, 8)), int(rng.integers(-40, 200))) for _ in range(40)]
    masks, bricks = ref_terrain(pos)
    some = 0
    for i, p in enumerate(pos):
        r = terrain.generate_sector_fastnoise(*p)
        if masks[i] == 0:
            assert r is None or int(r[0]) == 0, p
        else:
            _assert_sector(r[0], r[1], masks[i], bricks[i], p)
            some += 1
    assert some >= 10
    if not (Path(terrain.__file__).resolve().parent / "_ref" / "libterrain_gen.so").exists():
        pytest.skip("scenes/_ref/libterrain_gen.so not built")
    nx, ny, nz, shift = 3, 24, 2, 512
    big = terrain.terrain_fastnoise_big(nx, ny, nz, shift)
    pos = [(x, y, z) for y in range(ny) for z in range(nz) for x in range(nx)]
    masks, bricks = ref_terrain([(x, y - shift // 32, z) for x, y, z in pos])
    solid = 0
    for i, p in enumerate(pos):
        if masks[i] == 0:
            assert p not in big["sectors"], p
            continue
        m, b = big["sectors"][p]
        _assert_sector(m, b, masks[i], bricks[i], p)
        solid += int(masks[i]) == (1 << 64) - 1
    assert solid >= nx * nz * 10  # the rock under the surface: full sectors
