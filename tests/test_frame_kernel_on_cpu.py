"""The per-pixel body of the FRAME kernels in the CPU tier: voxelrt_b200/csrc/vrt_shade.cuh (primary_ray with the unguarded rcp / sqrt forms,
cast_ray, shade_pixel_primary for bounces = 0, shade_pixel with blue-noise bounces and the sky cube, G-buffer packing) compiled for the
host and run pixel by pixel with the product's own fill_frame_params — against the oracle's RenderRow restatement, byte for byte
(16 B/px tile framebuffer) and hit record for hit record.  Hardware behaviour and the warp-level forms (wavefront passes, CTA compaction,
macro steps) remain the GPU tests' business."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import assert_hits_equal
from test_glsl_kernel_on_cpu import DeviceLayout, EmuScene

NATIVE = Path(__file__).resolve().parent / "native"


@pytest.fixture(scope="module")
def emu():
    from voxelrt_b200 import capi

    subprocess.run(["make", "-s", "-C", str(NATIVE), "libemu_render.so"], check=True)
    lib = C.CDLL(str(NATIVE / "libemu_render.so"))
    lib.emu_render.argtypes = [C.POINTER(EmuScene), C.POINTER(capi.VrtFrame), C.c_void_p, C.c_void_p, C.POINTER(capi.VrtSkyDesc), C.c_void_p, C.c_void_p]
    lib.emu_render.restype = None
    return lib


@pytest.fixture(scope="module")
def layout(hash_scene):
    return DeviceLayout(hash_scene)


def _frames(w, h, bounces, frame_nos=(1,)):
    from scenes import camera
    from voxelrt_b200 import capi

    cams = [camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45), camera.Camera(pos=(20.5, 70.1, 150.25), yaw=2.4, pitch=-0.3),
            camera.Camera(pos=(100.0, 200.0, 100.0), yaw=1.0, pitch=-1.2)]
    for cam in cams:
        for fno in frame_nos:
            proj, inv, wo, frac = cam.matrices(w, h)
            yield capi.make_frame(w, h, inv, proj, wo, frac, frame_no=fno, bounces=bounces)


@pytest.mark.parametrize("bounces", [0, 1, 2])
def test_frame_kernel_source_equals_oracle(emu, layout, hash_oracle, shading_inputs, bounces):
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    hash_oracle.set_blue_noise(bn)
    hash_oracle.set_sky(desc, tex)
    bn_a = np.ascontiguousarray(bn, np.uint8)
    tex_a = np.ascontiguousarray(tex, np.uint32)
    w, h = 320, 180
    n_hits = 0
    for frame in _frames(w, h, bounces, frame_nos=(1, 2, 77) if bounces else (1,)):
        got = np.zeros(w * h // 16, capi.TILE_DTYPE)
        aux = np.zeros(w * h, capi.HIT_DTYPE)
        emu.emu_render(C.byref(layout.c), C.byref(frame), bn_a.ctypes.data, tex_a.ctypes.data, C.byref(desc), got.ctypes.data, aux.ctypes.data)
        f2 = capi.make_frame(w, h, list(frame.inv_proj), list(frame.proj), list(frame.world_origin), list(frame.origin_frac), frame_no=frame.frame_no, bounces=bounces)
        want, want_aux, _ = hash_oracle.render(f2, want_aux=True)
        for field in ("albedo", "depth", "irr_rg", "irr_bx"):
            a, b = got[field].view(np.uint32), want[field].view(np.uint32)
            assert np.array_equal(a, b), f"bounces={bounces} frame_no={frame.frame_no}: {field} differs at {(a != b).sum()} pixels"
        assert_hits_equal(aux, want_aux, f"bounces={bounces}: aux hit records")
        n_hits += int(((want_aux["flags"] & 0x100) != 0).sum())
    assert n_hits > 10000


def test_bench_terrain_frame_reference_camera(emu, bench_scene, bench_oracle):
    """BASELINE configs[0]/[1]'s scene (the FastNoise2 terrain, or the hash terrain over the same sectors where the reference's
    FastNoise2 is not in the tree) and the reference's camera (Main.cpp:76-78), primary rays: the frame kernel's pixel source
    reproduces the oracle's G-buffer bytes and hit records at 640x360."""
    from scenes import camera
    from voxelrt_b200 import capi

    L = DeviceLayout(bench_scene)
    w, h = 640, 360
    proj, inv, wo, frac = camera.Camera().matrices(w, h)
    frame = capi.make_frame(w, h, inv, proj, wo, frac, frame_no=1, bounces=0)
    got = np.zeros(w * h // 16, capi.TILE_DTYPE)
    aux = np.zeros(w * h, capi.HIT_DTYPE)
    emu.emu_render(C.byref(L.c), C.byref(frame), None, None, None, got.ctypes.data, aux.ctypes.data)
    want, want_aux, st = bench_oracle.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=1, bounces=0), want_aux=True)
    assert got.tobytes() == want.tobytes()
    assert_hits_equal(aux, want_aux, "bench terrain, reference camera")
    hits = int(((want_aux["flags"] & 0x100) != 0).sum())
    assert 0.6 < hits / (w * h) < 0.9  # SURVEY probe: 75 % of the primary rays hit


def test_sponza_frames_one_bounce(emu, shading_inputs):
    """BASELINE configs[2]'s scene (the bundled Sponza model voxelised into 1024^3, view 32x16x32 sectors), one bounce of blue-noise diffuse
    rays inside the atrium: the frame kernel's pixel source against the oracle at 384x216, two cameras and frame numbers."""
    from oracle import pyoracle
    from scenes import camera, models, terrain
    from voxelrt_b200 import capi

    if not models.sponza_available(1024):
        pytest.skip("scenes/_ref/sponza_1024.dat absent (built by __graft_entry__.build() where the reference assets exist)")
    scene = models.sponza(1024)
    (bn, _), (desc, tex, _) = shading_inputs
    orc = pyoracle.OracleMap(5, 4)
    orc.set_palette(scene["palette"])
    orc.sync(terrain.scene_records(scene))
    orc.set_blue_noise(bn)
    orc.set_sky(desc, tex)
    L = DeviceLayout(scene, sxz=5, sy=4)
    bn_a, tex_a = np.ascontiguousarray(bn, np.uint8), np.ascontiguousarray(tex, np.uint32)
    w, h = 384, 216
    cams = [camera.Camera(pos=(210.3, 80.2, 505.7), yaw=1.5, pitch=-0.15), camera.Camera(pos=(700.1, 300.4, 520.2), yaw=-1.2, pitch=-0.6)]
    for frame_no, cam in ((1, cams[0]), (3, cams[1])):
        proj, inv, wo, frac = cam.matrices(w, h)
        frame = capi.make_frame(w, h, inv, proj, wo, frac, frame_no=frame_no, bounces=1)
        got = np.zeros(w * h // 16, capi.TILE_DTYPE)
        aux = np.zeros(w * h, capi.HIT_DTYPE)
        emu.emu_render(C.byref(L.c), C.byref(frame), bn_a.ctypes.data, tex_a.ctypes.data, C.byref(desc), got.ctypes.data, aux.ctypes.data)
        want, want_aux, st = orc.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=frame_no, bounces=1), want_aux=True)
        assert st.hits > w * h
        assert got.tobytes() == want.tobytes(), f"sponza frame {frame_no}"
        assert_hits_equal(aux, want_aux, f"sponza frame {frame_no}")


@pytest.fixture(scope="module")
def kern():
    from voxelrt_b200 import capi

    subprocess.run(["make", "-s", "-C", str(NATIVE), "libemu_kernels.so"], check=True)
    lib = C.CDLL(str(NATIVE / "libemu_kernels.so"))
    lib.emu_render_kernel.argtypes = [C.POINTER(EmuScene), C.POINTER(capi.VrtFrame), C.c_void_p, C.c_void_p, C.POINTER(capi.VrtSkyDesc), C.c_void_p, C.c_uint32, C.c_uint32]
    lib.emu_render_kernel.restype = C.c_int
    return lib


@pytest.mark.parametrize("size", [(320, 180), (36, 4), (260, 148)])
def test_k_render_grid_and_the_screen_split(kern, layout, hash_oracle, shading_inputs, size):
    """k_render as launched (fill_frame_params + fill_frame_partition, warp tiles numbered inside 32x32 macro tiles or 8-pixel bands, store_pixel):
    the whole frame equals the oracle; for N = 2, 3, 8 ranks — by macro tile and by band (VRT_FRAME_PART_ROWS) — the parts are disjoint and
    their union is the whole frame, byte for byte (N-GPU == 1-GPU); row ranges (the band-pipelined host path) likewise."""
    from scenes import camera
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    hash_oracle.set_blue_noise(bn)
    hash_oracle.set_sky(desc, tex)
    bn_a, tex_a = np.ascontiguousarray(bn, np.uint8), np.ascontiguousarray(tex, np.uint32)
    w, h = size
    cam = camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45)
    proj, inv, wo, frac = cam.matrices(w, h)
    SENT = 0xA5A5A5A5

    def run(bounces, part_index=0, part_count=1, flags=0, row0=0, row1=0, into=None):
        fr = capi.make_frame(w, h, inv, proj, wo, frac, frame_no=5, bounces=bounces, flags=flags, part_index=part_index, part_count=part_count)
        out = into if into is not None else np.full(w * h * 4, SENT, np.uint32)
        assert kern.emu_render_kernel(C.byref(layout.c), C.byref(fr), bn_a.ctypes.data, tex_a.ctypes.data, C.byref(desc), out.ctypes.data, row0, row1) >= 0
        return out

    for bounces in (0, 1):
        want = hash_oracle.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=5, bounces=bounces))[0]
        want_u = np.frombuffer(want.tobytes(), np.uint32)
        whole = run(bounces)
        assert np.array_equal(whole, want_u), f"bounces={bounces}"
        for flags in (0, capi.VRT_FRAME_PART_ROWS):
            for n in (2, 3, 8):
                covered = np.zeros(w * h * 4, bool)
                acc = np.full(w * h * 4, SENT, np.uint32)
                for r in range(n):
                    part = run(bounces, r, n, flags)
                    mine = part != SENT
                    assert not (covered & mine).any(), f"parts overlap (n={n}, flags={flags})"
                    covered |= mine
                    acc[mine] = part[mine]
                assert np.array_equal(acc, want_u), f"union of {n} parts (flags={flags}, bounces={bounces})"
    # row ranges of an unpartitioned frame, written into one buffer
    acc = np.full(w * h * 4, SENT, np.uint32)
    rows_total = (h + 31) // 32
    for r0 in range(0, rows_total, 2):
        run(0, row0=r0, row1=min(r0 + 2, rows_total), into=acc)
    assert np.array_equal(acc, np.frombuffer(hash_oracle.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=5, bounces=0))[0].tobytes(), np.uint32))


def test_k_render_linear_output(kern, layout, hash_oracle):
    """VRT_FRAME_LINEAR_OUTPUT: four planes (albedo, depth, irrRG, irrBX) of w*h words instead of 4x4 tiles — same values."""
    from scenes import camera
    from voxelrt_b200 import capi

    w, h = 132, 68
    cam = camera.Camera(pos=(20.5, 70.1, 150.25), yaw=2.4, pitch=-0.3)
    proj, inv, wo, frac = cam.matrices(w, h)
    fr = capi.make_frame(w, h, inv, proj, wo, frac, frame_no=2, bounces=0, flags=capi.VRT_FRAME_LINEAR_OUTPUT)
    out = np.zeros(4 * w * h, np.uint32)
    assert kern.emu_render_kernel(C.byref(layout.c), C.byref(fr), None, None, None, out.ctypes.data, 0, 0) > 0
    want_lin = hash_oracle.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=2, bounces=0, flags=capi.VRT_FRAME_LINEAR_OUTPUT))[0]
    assert np.array_equal(out, np.ascontiguousarray(want_lin).view(np.uint32).reshape(-1))
    want_tiles = hash_oracle.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=2, bounces=0))[0]
    t = want_tiles.reshape(h // 4, w // 4)
    for k, name in enumerate(("albedo", "depth", "irr_rg", "irr_bx")):
        plane = t[name].view(np.uint32).reshape(h // 4, w // 4, 4, 4).transpose(0, 2, 1, 3).reshape(h, w)
        assert np.array_equal(out[k * w * h : (k + 1) * w * h].reshape(h, w), plane), name


def test_big_view_two_bounces_occ_form(emu, kern, hash_scene, shading_inputs):
    """BASELINE configs[3]'s shape in small: a 4096x512x4096 view (128x16x128 sectors), two bounces — the frame kernels of big views consult the
    one-bit sector table before the header (OCC instantiation).  Pixel source against the oracle, table from the emulated k_build_occ."""
    from oracle import pyoracle
    from scenes import camera, terrain
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    orc = pyoracle.OracleMap(7, 4)
    orc.set_palette(hash_scene["palette"])
    orc.sync(terrain.scene_records(hash_scene))
    orc.set_blue_noise(bn)
    orc.set_sky(desc, tex)
    L = DeviceLayout(hash_scene, sxz=7, sy=4)
    kern.emu_build_occ.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    kern.emu_build_occ.restype = None
    n_all = len(L.hdr_all)
    occ = np.zeros((n_all + 31) // 32 + 1, np.uint32)
    kern.emu_build_occ(L.hdr_all.ctypes.data, n_all, occ.ctypes.data)
    emu.emu_render_set_occ.argtypes = [C.c_void_p]
    emu.emu_render_set_occ.restype = None
    emu.emu_render_set_occ(occ.ctypes.data)
    try:
        bn_a, tex_a = np.ascontiguousarray(bn, np.uint8), np.ascontiguousarray(tex, np.uint32)
        w, h = 320, 180
        for cam in (camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45), camera.Camera(pos=(400.5, 260.1, 380.25), yaw=-2.3, pitch=-0.55)):
            proj, inv, wo, frac = cam.matrices(w, h)
            frame = capi.make_frame(w, h, inv, proj, wo, frac, frame_no=9, bounces=2)
            got = np.zeros(w * h // 16, capi.TILE_DTYPE)
            aux = np.zeros(w * h, capi.HIT_DTYPE)
            emu.emu_render(C.byref(L.c), C.byref(frame), bn_a.ctypes.data, tex_a.ctypes.data, C.byref(desc), got.ctypes.data, aux.ctypes.data)
            want, want_aux, _ = orc.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=9, bounces=2), want_aux=True)
            assert got.tobytes() == want.tobytes()
            assert_hits_equal(aux, want_aux, "big view, OCC form")
    finally:
        emu.emu_render_set_occ(None)


@pytest.mark.parametrize("size,bounces,flags_parts", [((256, 144), 1, (0, 1)), ((256, 144), 3, (0, 1)), ((132, 76), 2, (8, 3)), ((36, 8), 2, (0, 2))])
def test_wavefront_pipeline_source_equals_oracle(kern, layout, hash_oracle, shading_inputs, size, bounces, flags_parts):
    """The WAVEFRONT form of a frame with bounces as launched (k_wave_primary, then per level k_wave_trace — persistent warps whose lanes are
    refilled from the level's queue, the branch-lean trip, the one-trip shortcut for NaN directions — and k_wave_shade with the packet votes),
    all warps in lockstep on the host: the frame equals the oracle (= the reference's RenderRow) byte for byte, hit records included; with a
    screen split the parts are disjoint and their union is the frame."""
    from scenes import camera
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    hash_oracle.set_blue_noise(bn)
    hash_oracle.set_sky(desc, tex)
    bn_a, tex_a = np.ascontiguousarray(bn, np.uint8), np.ascontiguousarray(tex, np.uint32)
    kern.emu_wave_frame.argtypes = [C.POINTER(EmuScene), C.POINTER(capi.VrtFrame), C.c_void_p, C.c_void_p, C.POINTER(capi.VrtSkyDesc), C.c_void_p, C.c_void_p, C.c_uint32]
    kern.emu_wave_frame.restype = C.c_int
    w, h = size
    flags, parts = flags_parts
    SENT = 0xA5A5A5A5
    cams = [camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45), camera.Camera(pos=(60.3, 20.2, 40.7), yaw=1.2, pitch=-0.3),
            camera.Camera(pos=(96.3, 126.9, 20.7), yaw=0.2, pitch=1.2)]
    for k, cam in enumerate(cams):
        proj, inv, wo, frac = cam.matrices(w, h)
        want, want_aux, _ = hash_oracle.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=3 + 64 * k, bounces=bounces), want_aux=True)
        acc = np.full(w * h * 4, SENT, np.uint32)
        aux = np.zeros(w * h, capi.HIT_DTYPE)
        queued = 0
        for r in range(parts):
            fr = capi.make_frame(w, h, inv, proj, wo, frac, frame_no=3 + 64 * k, bounces=bounces, flags=flags, part_index=r, part_count=parts)
            part = np.full(w * h * 4, SENT, np.uint32)
            n1 = kern.emu_wave_frame(C.byref(layout.c), C.byref(fr), bn_a.ctypes.data, tex_a.ctypes.data, C.byref(desc), part.ctypes.data, aux.ctypes.data, 5)
            assert n1 >= 0
            queued += n1
            mine = part != SENT
            assert not ((acc != SENT) & mine).any()
            acc[mine] = part[mine]
        assert np.array_equal(acc, np.frombuffer(want.tobytes(), np.uint32)), f"camera {k}: {(acc != np.frombuffer(want.tobytes(), np.uint32)).sum()} words differ"
        assert_hits_equal(aux, want_aux, f"camera {k}: aux hit records")
        if k == 0:
            assert queued > w * h // 4  # (bounce rays really went through the queues)
