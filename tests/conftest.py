"""Shared fixtures.  `-m "not gpu"` runs here (no GPU): oracle vs golden vectors and independent restatements, host logic, ABI
exports, and the kernels' own source compiled for the host (tests/native/emu_*.cpp behind cuda_host_shim.h; test_*_on_cpu.py).  `-m gpu` runs on a B200: the parity tests proper, every one of them through the C ABI of
libvoxelrt_b200.so (voxelrt_b200.capi is a 1:1 ctypes binding of include/voxelrt_b200.h)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: more than a few seconds on CPU")


def _cuda_available() -> bool:
    import os

    if os.environ.get("VRT_ASSUME_CUDA") == "1":  # quick GPU sessions: skip the (slow on a fresh box) torch import
        return True
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a GPU test on a box without a GPU must FAIL loudly only when explicitly selected (-m gpu);
    # in a plain `pytest tests/` run on a CPU box they are skipped.
    if _cuda_available():
        return
    selected_gpu = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    if selected_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def hash_scene():
    from scenes import terrain

    return terrain.terrain_hash(6, 4, 6, seed=77)


@pytest.fixture(scope="session")
def hash_oracle(hash_scene):
    from oracle import pyoracle
    from scenes import terrain

    orc = pyoracle.OracleMap(6, 4)
    orc.set_palette(hash_scene["palette"])
    orc.sync(terrain.scene_records(hash_scene))
    return orc


@pytest.fixture(scope="session")
def bench_scene():
    """The scene of BASELINE.json configs 1/2 (FastNoise2 terrain when scenes/_ref holds the library)."""
    from scenes import terrain

    return terrain.bench_terrain()


@pytest.fixture(scope="session")
def bench_oracle(bench_scene):
    from oracle import pyoracle
    from scenes import terrain

    orc = pyoracle.OracleMap(6, 4)
    orc.set_palette(bench_scene["palette"])
    orc.sync(terrain.scene_records(bench_scene))
    return orc


@pytest.fixture(scope="session")
def shading_inputs():
    from scenes import shading

    return shading.load_blue_noise(), shading.load_sky()


def ctx_for(scene, xz=6, y=4, **kw):
    from scenes import terrain
    from voxelrt_b200 import capi

    ctx = capi.Context(xz, y, device=0, **kw)
    ctx.set_palette(scene["palette"])
    ctx.sync(terrain.scene_records(scene))
    return ctx


@pytest.fixture(scope="session")
def hash_ctx(hash_scene):
    return ctx_for(hash_scene)


@pytest.fixture(scope="session")
def bench_ctx(bench_scene):
    return ctx_for(bench_scene)


def random_rays(rng, n, extent_xz, extent_y, wo):
    """Rays starting inside / around the view box with random directions (float32)."""
    o = np.stack(
        [rng.uniform(-8, extent_xz + 8, n), rng.uniform(-8, extent_y + 8, n), rng.uniform(-8, extent_xz + 8, n)], axis=1
    )
    o = (o - np.asarray(wo, np.float64)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def assert_hits_equal(a, b, what="", ignore_iters=False):
    """Bit-exact comparison of two VrtHit arrays: every field, float fields by bit pattern (so -0.0
    != +0.0); the only tolerance is the NaN payload (x86 produces the default NaN 0xFFC00000, the GPU
    0x7FFFFFFF — both are "NaN" to every consumer)."""
    for name in a.dtype.names:
        x, y = a[name], b[name]
        if x.dtype.kind == "f":
            both_nan = np.isnan(x) & np.isnan(y)
            x, y = x.view(np.uint32), y.view(np.uint32)
            bad = np.nonzero((x != y) & ~both_nan)[0]
        else:
            if name == "flags" and ignore_iters:
                x, y = x & 0xFFFF, y & 0xFFFF
            bad = np.nonzero(x != y)[0]
        assert bad.size == 0, f"{what}: field {name} differs at {bad.size} rays, first {bad[:5]}: {a[bad[:3]]} vs {b[bad[:3]]}"
