"""Runs the C++ host tests (tests/native/test_host.cpp): VoxelMap mirror, indexers, slot arena; with a
GPU also the B200Renderer adapter end to end (frames after brush-like edits == oracle rebuilt from scratch)."""
from __future__ import annotations

import subprocess
from pathlib import Path

import pytest

EXE = Path(__file__).resolve().parent / "native" / "test_host"


def _run(*args):
    if not EXE.exists():
        subprocess.run(["make", "-s", "-C", str(EXE.parent)], check=True)
    r = subprocess.run([str(EXE), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "native tests: ok" in r.stdout, r.stdout + r.stderr
    return r.stdout


def test_host_logic_cpu():
    _run()


@pytest.mark.gpu
def test_adapter_end_to_end_gpu():
    out = _run("--gpu")
    assert "gpu:" in out


def test_adapter_inverse_proj_screen_matches_reference_and_python():
    """GetInverseProjScreenMat of the C++ adapter == the reference's own GBuffer::GetInverseProjScreenMat (compiled into oracle/_ref over
    the stand-in glm, which restates GLM's float cofactor inverse) == scenes/camera.py, bit for bit (the adapter used a
    double-precision Gauss-Jordan inverse, so B200Renderer::RenderFrame generated other primary rays than the tests fed)."""
    import ctypes as C

    import numpy as np

    from oracle import refharness
    from scenes import camera

    lib = C.CDLL(str(Path(__file__).resolve().parents[1] / "voxelrt_b200" / "lib" / "libvoxelrt_b200_host.so"))
    lib.vrt_host_inverse_proj_screen.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.vrt_host_inverse_proj_screen.restype = None
    rng = np.random.default_rng(3)
    have_ref = refharness.available()
    for i in range(64):
        cam = camera.Camera(pos=rng.random(3) * 900.0, yaw=rng.random() * 6.28, pitch=rng.random() * 2.4 - 1.2, fov_deg=40.0 + 80.0 * rng.random())
        w, h = [(1280, 720), (3840, 2160), (1920, 1080), (260, 148)][i % 4]
        pv, inv, _, _ = cam.matrices(w, h)
        out = np.zeros(16, np.float32)
        lib.vrt_host_inverse_proj_screen(pv.ctypes.data, w, h, out.ctypes.data)
        assert np.array_equal(out.view(np.uint32), inv.view(np.uint32)), i
        if have_ref:
            assert np.array_equal(refharness.inverse_proj_screen(pv, w, h).view(np.uint32), inv.view(np.uint32)), i
