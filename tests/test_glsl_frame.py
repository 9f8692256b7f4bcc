"""VRT_FRAME_GLSL frames — the frame shader of the reference's GPU renderer (Shaders/VoxelRender.comp:29-93) per pixel.

PARITY UNPINNED: GLSL needs a GL device, so nothing here is compared with the reference's own output.  What is checked:
  * the oracle (orc_render_glsl) against an INDEPENDENT restatement of main() in plain Python over the independent cast model of
    tests/glsl_cast_model_py.py, pixel by pixel, bit for bit (CPU tier);
  * the kernel's source (vrt_glsl_frame.cuh compiled for the host) against the oracle on whole frames (CPU tier);
  * the CUDA kernel against the oracle through the C ABI, every way a frame can be delivered (GPU tier).
The canonical arithmetic and the two stand-ins (sky sampler, unassigned `out` fields) are stated in vrt_glsl_frame.cuh."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from test_glsl_kernel_on_cpu import DeviceLayout, EmuScene

NATIVE = Path(__file__).resolve().parent / "native"
F = np.float32


def _cams():
    from scenes import camera

    return [camera.Camera(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45), camera.Camera(pos=(20.5, 70.1, 150.25), yaw=2.4, pitch=-0.3),
            camera.Camera(pos=(100.0, 40.0, 100.0), yaw=1.0, pitch=-0.2)]


def _frame(cam, w, h, bounces, frame_no=1, aniso=False, **kw):
    from voxelrt_b200 import capi

    proj, inv, wo, frac = cam.matrices(w, h)
    flags = capi.VRT_FRAME_GLSL | (capi.VRT_FRAME_GLSL_ANISOTROPIC if aniso else 0) | kw.pop("flags", 0)
    return capi.make_frame(w, h, inv, proj, wo, frac, frame_no=frame_no, bounces=bounces, flags=flags, **kw)


@pytest.fixture(scope="module")
def shaded_oracle(hash_oracle, shading_inputs):
    (bn, _), (desc, tex, _) = shading_inputs
    hash_oracle.set_blue_noise(bn)
    hash_oracle.set_sky(desc, tex)
    return hash_oracle


# ---------------------------------------------------------------------------------------------------------------------------------
# an independent restatement of main() (VoxelRender.comp:29-93), numpy float32 scalars, one operation at a time
# ---------------------------------------------------------------------------------------------------------------------------------
def _normalize(v):
    k = F(1) / np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2], dtype=F)
    return [v[0] * k, v[1] * k, v[2] * k]


def _mat_color(md):
    c = [F((md >> 11) & 31) * (F(1) / F(31)), F((md >> 5) & 63) * (F(1) / F(63)), F(md & 31) * (F(1) / F(31))]
    return [x * x for x in c]


def _emission(md):
    return F(np.array([md >> 16], np.uint16).view(np.float16)[0])


def _half_bits(x):
    return int(np.array([x], F).astype(np.float16).view(np.uint16)[0])


def _unorm8(v):
    return int(np.rint(min(max(v, F(0)), F(1)) * F(255)))


class _Model:
    def __init__(self, scene, oracle, bn):
        from glsl_cast_model_py import DenseWorld

        self.world = DenseWorld(scene)
        self.world.palette = np.asarray(scene["palette"], np.uint64)
        self.oracle = oracle
        self.bn = np.asarray(bn, np.uint8).reshape(-1, 128, 2)  # [y (64 slices of 128), x, (r, g)]

    def cast(self, o, d, wo, coarse, aniso, hit):
        from glsl_cast_model_py import cast

        r = cast(self.world, o, d, wo, coarse_mode=coarse, aniso=aniso)
        if not r["capped"]:
            hit["iters"] = r["iters"]
        if not r["hit"]:
            return False
        hit["mat"], hit["pos"], hit["nrm"] = r["material"], list(r["pos"]), [F(n) for n in r["normal"]]
        return True

    def sky(self, d):
        t = self.oracle.sky_sample(np.array(d, F), 0)  # the CPU renderer's cube at level 0 (pinned against the reference), texel * 3
        return [min((F(v) / F(3)) * F(5), F(50000)) for v in t]

    def random_dir(self, x, y, frame_no, i):
        fi = F(i)
        sx, sy = fi * F(0.75487766624669276005) + F(0.5), fi * F(0.56984029099805326591) + F(0.5)
        sx, sy = sx - np.floor(sx), sy - np.floor(sy)
        px, py = (x + int(sx * F(128))) & 127, ((y + int(sy * F(128))) & 127) + (frame_no & 63) * 128
        t = self.bn[py, px]
        nx, ny = (F(t[0]) + F(0.5)) * F(1 / 256), (F(t[1]) + F(0.5)) * F(1 / 256)
        yy = nx * F(2) - F(1)
        s, c = self.oracle.sincos_2pi(float(ny))
        sq = np.sqrt(F(1) - yy * yy, dtype=F)
        return [F(s) * sq, yy, F(c) * sq]

    def pixel(self, f, x, y):
        with np.errstate(all="ignore"):
            iv, pj = [F(v) for v in f.inv_proj], [F(v) for v in f.proj]
            wo = [int(v) for v in f.world_origin]
            fx, fy = F(x), F(y)
            nr = [((iv[k] * fx + iv[4 + k] * fy) + iv[8 + k] * F(0)) + iv[12 + k] * F(1) for k in range(4)]
            fr = [nr[k] + iv[8 + k] for k in range(4)]
            i_n, i_f = F(1) / nr[3], F(1) / fr[3]
            pos = [nr[a] * i_n + F(f.origin_frac[a]) for a in range(3)]
            d = _normalize([fr[a] * i_f for a in range(3)])
            aniso = bool(f.flags & 32)
            hit = {"iters": 0, "mat": 0, "pos": [F(0)] * 3, "nrm": [F(0)] * 3}
            nrm, depth = [F(0)] * 3, F(-1)
            if self.cast(pos, d, wo, False, aniso, hit):
                albedo = _mat_color(hit["mat"])
                nrm = list(hit["nrm"])
                hp = [v * F(0.0625) for v in hit["pos"]]
                pz = ((pj[2] * hp[0] + pj[6] * hp[1]) + pj[10] * hp[2]) + pj[14] * F(1)
                pw = ((pj[3] * hp[0] + pj[7] * hp[1]) + pj[11] * hp[2]) + pj[15] * F(1)
                depth = pz / pw
                em = _emission(hit["mat"])
                irr = list(albedo) if f.bounces == 0 else [a * em for a in albedo]
                thr = [F(1)] * 3
                sun = _normalize([F(0.3), F(0.9), F(-0.28)])
                sun_col = [F(1.2) * F(5), F(1.1) * F(5), F(1.0) * F(5)]
                scratch = dict(hit)
                if f.bounces:
                    so = [hit["pos"][a] + hit["nrm"][a] * F(0.01) for a in range(3)]
                    if not self.cast(so, sun, wo, True, aniso, scratch):
                        irr = [irr[a] + sun_col[a] for a in range(3)]
                    else:
                        thr = [t * F(0.5) for t in thr]
                for i in range(f.bounces):
                    rnd = self.random_dir(x, y, f.frame_no, i)
                    pos = [hit["pos"][a] + hit["nrm"][a] * F(0.01) for a in range(3)]
                    d = _normalize([hit["nrm"][a] + rnd[a] for a in range(3)])
                    if not self.cast(pos, d, wo, True, aniso, hit):
                        sky = self.sky(d)
                        irr = [irr[a] + thr[a] * sky[a] for a in range(3)]
                        break
                    col = _mat_color(hit["mat"])
                    thr = [thr[a] * col[a] for a in range(3)]
                    em = _emission(hit["mat"])
                    if i < 2:
                        so = [hit["pos"][a] + hit["nrm"][a] * F(0.01) for a in range(3)]
                        if not self.cast(so, sun, wo, True, aniso, scratch):
                            thr = [thr[a] * sun_col[a] for a in range(3)]
                            em = em + F(5)
                    irr = [irr[a] + thr[a] * em for a in range(3)]
            else:
                irr, albedo = self.sky(d), [F(1)] * 3
            code = sum(int(min(max(nrm[a] + F(1), F(0)), F(3))) << (2 * a) for a in range(3))
            alb = _unorm8(albedo[0]) | _unorm8(albedo[1]) << 8 | _unorm8(albedo[2]) << 16 | code << 24
            return alb, np.array([depth], F).view(np.uint32)[0], _half_bits(irr[0]) | _half_bits(irr[1]) << 16, _half_bits(irr[2]) | _half_bits(F(hit["iters"])) << 16


def _pixel_of(tiles, w, x, y):
    t = tiles[(y >> 2) * (w >> 2) + (x >> 2)]
    l = (x & 3) | ((y & 3) << 2)
    return int(t["albedo"][l]), int(t["depth"].view(np.uint32)[l]), int(t["irr_rg"][l]), int(t["irr_bx"][l])


@pytest.mark.parametrize("bounces,aniso", [(0, False), (2, False), (3, True)])
def test_oracle_equals_the_independent_python_restatement(shaded_oracle, hash_scene, shading_inputs, bounces, aniso):
    (bn, _), _ = shading_inputs
    model = _Model(hash_scene, shaded_oracle, bn)
    w, h = 64, 36
    rng = np.random.default_rng(bounces)
    n_hit = n_lit = 0
    for k, cam in enumerate(_cams()):
        f = _frame(cam, w, h, bounces, frame_no=3 + 64 * k, aniso=aniso)
        want, _ = shaded_oracle.render_glsl(f)
        for _ in range(24):
            x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
            got = model.pixel(f, x, y)
            assert got == _pixel_of(want, w, x, y), (k, x, y)
            n_hit += got[1] != np.array([-1.0], F).view(np.uint32)[0]
            n_lit += (got[2] & 0xFFFF) != 0
    assert n_hit >= 30 and n_lit >= 20


# ---------------------------------------------------------------------------------------------------------------------------------
# the kernel source on the CPU
# ---------------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu():
    from voxelrt_b200 import capi

    subprocess.run(["make", "-s", "-C", str(NATIVE), "libemu_kernels.so"], check=True)
    lib = C.CDLL(str(NATIVE / "libemu_kernels.so"))
    lib.emu_render_glsl.argtypes = [C.POINTER(EmuScene), C.POINTER(capi.VrtFrame), C.c_void_p, C.c_void_p, C.POINTER(capi.VrtSkyDesc), C.c_void_p, C.c_void_p]
    lib.emu_render_glsl.restype = C.c_int
    return lib


@pytest.mark.parametrize("bounces,aniso", [(0, False), (1, False), (3, False), (2, True)])
def test_kernel_source_equals_oracle(emu, shaded_oracle, hash_scene, shading_inputs, bounces, aniso):
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    L = DeviceLayout(hash_scene)
    bn_a, tex_a = np.ascontiguousarray(bn, np.uint8), np.ascontiguousarray(tex, np.uint32)
    w, h = 256, 144
    lit = 0
    for k, cam in enumerate(_cams()):
        f = _frame(cam, w, h, bounces, frame_no=1 + 70 * k, aniso=aniso)
        got = np.zeros(w * h // 16, capi.TILE_DTYPE)
        assert emu.emu_render_glsl(C.byref(L.c), C.byref(f), bn_a.ctypes.data, tex_a.ctypes.data, C.byref(desc), L.lut.ctypes.data, got.ctypes.data) > 0
        want, st = shaded_oracle.render_glsl(f)
        for field in ("albedo", "depth", "irr_rg", "irr_bx"):
            a, b = got[field].view(np.uint32), want[field].view(np.uint32)
            assert np.array_equal(a, b), f"bounces={bounces} cam {k}: {field} differs at {(a != b).sum()} pixels"
        assert st.iters > w * h
        lit += int((want["depth"] != -1.0).sum())
    assert lit > w * h  # more than a frame's worth of pixels see geometry over the three cameras


def test_kernel_source_band_parts_tile_the_frame(emu, shaded_oracle, hash_scene, shading_inputs):
    """VRT_FRAME_PART_ROWS: 3 ranks' bands, each rendered by its own launch, add up to the unsplit frame."""
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    L = DeviceLayout(hash_scene)
    bn_a, tex_a = np.ascontiguousarray(bn, np.uint8), np.ascontiguousarray(tex, np.uint32)
    w, h = 128, 72
    cam = _cams()[0]
    want, _ = shaded_oracle.render_glsl(_frame(cam, w, h, 1, frame_no=9))
    got = np.zeros(w * h // 16, capi.TILE_DTYPE)
    for r in range(3):
        f = _frame(cam, w, h, 1, frame_no=9, flags=capi.VRT_FRAME_PART_ROWS, part_index=r, part_count=3)
        assert emu.emu_render_glsl(C.byref(L.c), C.byref(f), bn_a.ctypes.data, tex_a.ctypes.data, C.byref(desc), L.lut.ctypes.data, got.ctypes.data) >= 0
        part, _ = shaded_oracle.render_glsl(f)
        band = (np.arange(h // 4) * 4 // 8) % 3 == r  # tile rows of this rank
        rows = np.repeat(band, w // 4)
        assert part[rows].tobytes() == want[rows].tobytes() and not part[~rows].view(np.uint8).any()
    assert got.tobytes() == want.tobytes()


# ---------------------------------------------------------------------------------------------------------------------------------
# the CUDA kernel through the C ABI
# ---------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("bounces,aniso", [(0, False), (1, False), (3, False), (2, True)])
def test_gpu_glsl_frames_equal_oracle(hash_ctx, shaded_oracle, shading_inputs, bounces, aniso):
    (bn, _), (desc, tex, _) = shading_inputs
    hash_ctx.set_blue_noise(bn)
    hash_ctx.set_sky(desc, tex)
    w, h = 640, 360
    for k, cam in enumerate(_cams()):
        f = _frame(cam, w, h, bounces, frame_no=1 + 70 * k, aniso=aniso)
        got, _ = hash_ctx.render(f)
        want, _ = shaded_oracle.render_glsl(_frame(cam, w, h, bounces, frame_no=1 + 70 * k, aniso=aniso))
        for field in ("albedo", "depth", "irr_rg", "irr_bx"):
            a, b = got[field].view(np.uint32), want[field].view(np.uint32)
            assert np.array_equal(a, b), f"bounces={bounces} cam {k}: {field} differs at {(a != b).sum()} pixels"


@pytest.mark.gpu
def test_gpu_glsl_frame_every_delivery(hash_ctx, shaded_oracle, shading_inputs):
    """Planes (VRT_FRAME_LINEAR_OUTPUT), device buffers on a caller stream and the band split of the host-buffer call all carry the oracle's
    bytes; flags the shader has no meaning for, and views without a 128^3 level, are refused."""
    import torch

    from scenes import terrain
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    hash_ctx.set_blue_noise(bn)
    hash_ctx.set_sky(desc, tex)
    w, h = 320, 184
    cam = _cams()[1]
    want, _ = shaded_oracle.render_glsl(_frame(cam, w, h, 2, frame_no=5))
    planes, _ = hash_ctx.render(_frame(cam, w, h, 2, frame_no=5, flags=capi.VRT_FRAME_LINEAR_OUTPUT))
    wp, _ = shaded_oracle.render_glsl(_frame(cam, w, h, 2, frame_no=5, flags=capi.VRT_FRAME_LINEAR_OUTPUT))
    assert planes.tobytes() == wp.tobytes()
    fb = torch.zeros(w * h * 4, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    hash_ctx.render_device(_frame(cam, w, h, 2, frame_no=5), fb.data_ptr(), None, st.cuda_stream)
    st.synchronize()
    assert fb.cpu().numpy().tobytes() == want.tobytes()
    host = np.zeros(w * h // 16, capi.TILE_DTYPE)
    for r in range(3):
        f = _frame(cam, w, h, 2, frame_no=5, flags=capi.VRT_FRAME_PART_ROWS, part_index=r, part_count=3)
        hash_ctx._chk(hash_ctx.lib.vrt_render(hash_ctx.h, C.byref(f), host.ctypes.data, None))
    assert host.tobytes() == want.tobytes()
    for bad in (capi.VRT_FRAME_COMPACT, capi.VRT_FRAME_AUX_HITS):
        f = _frame(cam, w, h, 0, flags=bad)
        aux = np.zeros(w * h, capi.HIT_DTYPE)
        assert hash_ctx.lib.vrt_render(hash_ctx.h, C.byref(f), host.ctypes.data, aux.ctypes.data) == capi.VRT_ERR_INVALID
    small = capi.Context(1, 1, device=0)
    f = _frame(cam, 64, 64, 0)
    assert small.lib.vrt_render(small.h, C.byref(f), host.ctypes.data, None) == capi.VRT_ERR_UNSUPPORTED  # no 128^3 level in a 2x2x2-sector view
    small.close()
