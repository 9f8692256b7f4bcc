"""The C-ABI library loads and exports every symbol include/voxelrt_b200.h declares; struct layouts of
the ctypes binding match the header; no compute is attempted (no GPU here)."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "voxelrt_b200.h").read_text()
    return sorted(set(re.findall(r"VRT_API\s+[\w\s\*]+?\b(vrt_\w+)\s*\(", text)))


def test_header_symbols_all_exported():
    from voxelrt_b200 import capi

    lib = capi.load()
    declared = _declared_symbols()
    assert len(declared) >= 19
    assert sorted(capi.EXPORTS) == declared, "capi.EXPORTS must list exactly what the header declares"
    for name in declared:
        assert hasattr(lib, name), f"libvoxelrt_b200.so does not export {name}"


def test_struct_sizes_match_header():
    from voxelrt_b200 import capi

    assert C.sizeof(capi.VrtConfig) == 24
    assert C.sizeof(capi.VrtDirtySector) == 40
    assert C.sizeof(capi.VrtFrame) == 8 + 64 + 64 + 12 + 12 + 16 + 8
    assert C.sizeof(capi.VrtSkyDesc) == 4 * 3 + 64 + 4 + 8  # padded to 8
    assert capi.HIT_DTYPE.itemsize == 48 and capi.HITD_DTYPE.itemsize == 48 and capi.TILE_DTYPE.itemsize == 256
    assert C.sizeof(capi.VrtStats) == 80 and C.sizeof(capi.VrtTraversalMetrics) == 48


def test_no_cpu_fallback_without_a_gpu():
    """On a machine without a CUDA device vrt_create must FAIL loudly (VRT_ERR_CUDA), never fall back."""
    import torch

    from voxelrt_b200 import capi

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.VrtError) as e:
        capi.Context(6, 4)
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)


def test_argument_validation_without_a_gpu():
    from voxelrt_b200 import capi

    lib = capi.load()
    h = C.c_void_p()
    assert lib.vrt_create(None, C.byref(h)) == -1
    bad = capi.VrtConfig(4, 0, 6, 4, 0, 0)  # wrong struct_size
    assert lib.vrt_create(C.byref(bad), C.byref(h)) == -1
    assert b"struct_size" in lib.vrt_last_error(None)
    assert lib.vrt_sync(None, 0, None) == -1 and lib.vrt_render(None, None, None, None) == -1
    assert lib.vrt_trace_glsl(None, 0, None, None, None, 0, None) == -1


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under voxelrt_b200/ may import, link or dlopen it."""
    for p in (ROOT / "voxelrt_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h", ".hpp") or p.name == "Makefile":
            t = p.read_text(errors="ignore")
            assert "pyoracle" not in t and "liboracle" not in t and "vrt_oracle" not in t and "refharness" not in t, p
            assert "pypostoracle" not in t and "post_oracle" not in t and "orc_trace" not in t, p


def test_post_header_symbols_all_exported():
    """include/voxelrt_b200_post.h (the GBuffer step) lives in the same shared library."""
    from voxelrt_b200 import post

    text = (ROOT / "include" / "voxelrt_b200_post.h").read_text()
    declared = sorted(set(re.findall(r"VRT_API\s+[\w\s\*]+?\b(vrt_gbuffer_\w+)\s*\(", text)))
    assert len(declared) == 11 and sorted(post.EXPORTS) == declared
    lib = post.load()
    for name in declared:
        assert hasattr(lib, name), f"libvoxelrt_b200.so does not export {name}"
    assert C.sizeof(post.VrtGBufferCamera) == 8 + 64 + 64 + 24 + 8 and post.RECORD_DTYPE.itemsize == 16


def test_gbuffer_has_no_cpu_fallback_and_validates_arguments():
    import torch

    from voxelrt_b200 import capi, post

    lib = post.load()
    assert lib.vrt_gbuffer_create(0, None) == -1
    assert lib.vrt_gbuffer_set_passes(None, 5) == -1 and lib.vrt_gbuffer_set_camera(None, None) == -1
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.VrtError) as e:
        post.GBuffer(0)
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)
