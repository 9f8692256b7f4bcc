"""THE TRAVERSAL's own source in the CPU tier: voxelrt_b200/csrc/vrt_device.cuh (cast_ray = the magic-number fast loop of the frame
kernels, or the generic loop for special rays; cast_finish; store_hit — exactly what k_trace runs per thread) compiled for the host
behind tests/native/cuda_host_shim.h and run lane by lane over a device-layout brickmap built with numpy — against the oracle AND
against the golden vectors the reference's own CpuRenderer.cpp produced (tests/golden/ref_trace_lane.npz).  The PTX helpers (LOP3
spellings, register pinning, rcp + Newton) have plain-C twins under VRT_HOST_EMULATION; macro steps (which need the box builder) are
off.  What this cannot see is the hardware: the GPU tests stay the parity tests proper."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import assert_hits_equal, random_rays
from test_glsl_kernel_on_cpu import DeviceLayout, EmuScene
from test_glsl_oracle import camera_frame_rays

NATIVE = Path(__file__).resolve().parent / "native"
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", str(NATIVE), "libemu_trace.so"], check=True)
    lib = C.CDLL(str(NATIVE / "libemu_trace.so"))
    lib.emu_trace.argtypes = [C.POINTER(EmuScene), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
    lib.emu_trace.restype = None
    return lib


@pytest.fixture(scope="module")
def layout(hash_scene):
    return DeviceLayout(hash_scene)


def _trace(emu, L, o, d, wo, max_iters=0, mode=0):
    from voxelrt_b200 import capi

    o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
    d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
    out = np.zeros(len(o), capi.HIT_DTYPE)
    w = (C.c_int32 * 3)(*[int(v) for v in wo])
    nf = C.c_uint64()
    emu.emu_trace(C.byref(L.c), w, o.ctypes.data, d.ctypes.data, max_iters, len(o), out.ctypes.data, mode, C.byref(nf))
    return out, int(nf.value)


def test_fast_loop_source_equals_oracle(emu, layout, hash_oracle):
    """Camera-frame and far-origin rays: nearly all take the fast loop; every VrtHit field equals the oracle's (which is pinned to the
    reference lane-wise)."""
    total = fast = 0
    for k in range(4):
        wo, o, d = camera_frame_rays(50000, 2000 + k)
        got, nf = _trace(emu, layout, o, d, wo)
        assert_hits_equal(got, hash_oracle.trace(o, d, wo)[0], f"camera frame wo={wo}")
        total, fast = total + len(o), fast + nf
    wo = (96, 64, 96)
    o, d = random_rays(np.random.default_rng(21), 100000, 192, 128, wo)
    got, nf = _trace(emu, layout, o, d, wo)
    assert_hits_equal(got, hash_oracle.trace(o, d, wo)[0], "far origins")
    total, fast = total + len(o), fast + nf
    assert fast > 0.99 * total, (fast, total)
    for cap in (1, 7, 300):  # iteration caps incl. the capped-ray bookkeeping
        got, _ = _trace(emu, layout, o[:20000], d[:20000], wo, max_iters=cap)
        assert_hits_equal(got, hash_oracle.trace(o[:20000], d[:20000], wo, max_iters=cap)[0], f"cap {cap}")


def test_generic_loop_source_equals_oracle(emu, layout, hash_oracle):
    """Zero / denormal / huge / NaN components and far-away world origins take the generic loop (x86 min / cvt semantics spelled out)."""
    vals = np.array([0.0, -0.0, 1.0, -1.0, 1e-39, -1e-39, 1e30, np.inf, -np.inf, np.nan, 0.3, -0.7], np.float32)
    d = np.array([(a, b, c) for a in vals for b in vals for c in vals], np.float32)
    o = np.tile(np.array([[0.25, 0.5, 0.75]], np.float32), (len(d), 1))
    o[::7, 0] = np.nan
    o[3::11, 2] = 1e30
    got, nf = _trace(emu, layout, o, d, (90, 70, 90))
    assert_hits_equal(got, hash_oracle.trace(o, d, (90, 70, 90))[0], "special directions")
    assert nf < len(o) // 2
    wo, o, d = camera_frame_rays(30000, 2100)
    got, _ = _trace(emu, layout, o, d, wo, mode=1)  # ordinary rays forced through the generic loop
    assert_hits_equal(got, hash_oracle.trace(o, d, wo)[0], "generic loop, ordinary rays")


def test_traversal_source_reproduces_the_reference_golden_vectors(emu, layout, hash_scene):
    """tests/golden/ref_trace_lane.npz holds RayCast of the reference's own CpuRenderer.cpp (compiled from /root/reference into
    oracle/_ref when the fixture was made), one active lane per packet.  The CUDA traversal source must reproduce it."""
    from scenes import terrain

    z = np.load(GOLD / "ref_trace_lane.npz")
    assert str(z["scene_digest"]) == terrain.scene_digest(hash_scene)
    want = z["hits"]
    got, nf = _trace(emu, layout, z["origin"], z["dir"], z["world_origin"])
    hit = (want["flags"] & 0x100) != 0
    assert np.array_equal(got["flags"] & 0x13F, want["flags"] & 0x13F)
    assert np.array_equal(got["material"], want["material"])
    for f in ("dist", "px", "py", "pz", "u", "v"):
        ok = (got[f].view(np.uint32) == want[f].view(np.uint32)) | (np.isnan(got[f]) & np.isnan(want[f]))
        assert ok.all(), f
    for f in ("vx", "vy", "vz"):
        assert np.array_equal(got[f][hit], want[f][hit]), f
    assert hit.sum() > 1000 and nf > 0


def test_traversal_counters_equal_the_oracle(emu, layout, hash_oracle):
    """The METRICS instantiation of the traversal (bench.py's roofline numerator: B = 8 I_s + 8 I_c + 9 H + 16 P is computed from these counters):
    iterations, sector-mask fetches, cell-mask fetches, hits and capped rays, summed over 150 k rays, equal the oracle's OrcStats."""
    from voxelrt_b200 import capi

    for k, cap in ((0, 0), (1, 40), (2, 0)):
        if k < 2:
            wo, o, d = camera_frame_rays(50000, 2500 + k)
        else:
            wo = (96, 64, 96)
            o, d = random_rays(np.random.default_rng(77), 50000, 192, 128, wo)
        out = np.zeros(len(o), capi.HIT_DTYPE)
        w = (C.c_int32 * 3)(*[int(v) for v in wo])
        words = (C.c_uint64 * 6)()
        emu.emu_trace(C.byref(layout.c), w, np.ascontiguousarray(o).ctypes.data, np.ascontiguousarray(d).ctypes.data, cap, len(o), out.ctypes.data, 4,
                      C.cast(words, C.POINTER(C.c_uint64)))
        want, st = hash_oracle.trace(o, d, wo, max_iters=cap)
        assert_hits_equal(out, want, f"metrics instantiation, case {k}")
        assert (int(words[1]), int(words[2]), int(words[3]), int(words[4]), int(words[5])) == (st.iters, st.sector_fetches, st.cell_fetches, st.hits, st.capped), (list(words), st.as_dict())
        assert st.algorithmic_bytes(0) == 8 * int(words[2]) + 8 * int(words[3]) + 9 * int(words[4])
