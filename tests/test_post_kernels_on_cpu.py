"""The GBuffer kernels' OWN SOURCE (the device half of voxelrt_b200/csrc/vrt_post.cu), compiled for the host behind
tests/native/cuda_host_shim.h and executed over the whole grid in a loop, against the image-space oracle — bit-exact, in the CPU
tier.  The kernels are plain per-thread functions (no shared memory, no barriers), so this runs the arithmetic the GPU runs; what it
cannot see is hardware behaviour (the GPU tests in test_gpu_zz_post.py do that).  The frame loop below restates the host side of
vrt_gbuffer_set_camera / vrt_gbuffer_denoise_present_device (buffer rotation of GBuffer.h:100-130)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from scenes import gbuffer_synth as pu

NATIVE = Path(__file__).resolve().parent / "native"


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", str(NATIVE), "libemu_post.so"], check=True)
    lib = C.CDLL(str(NATIVE / "libemu_post.so"))
    vp, i = C.c_void_p, C.c_int
    lib.emu_reproject.argtypes = [vp] * 11 + [i, i, i]
    lib.emu_variance.argtypes = [vp, vp, vp, i, i]
    lib.emu_atrous.argtypes = [vp, vp, i, i, i]
    lib.emu_present.argtypes = [vp, vp, vp, i, i, i]
    lib.emu_blit_only.argtypes = [vp, vp, vp, i, i]
    for f in (lib.emu_reproject, lib.emu_variance, lib.emu_atrous, lib.emu_present, lib.emu_blit_only):
        f.restype = None
    return lib


class EmulatedGBuffer:
    """Host side of VrtGBuffer over numpy buffers, kernels = the emulated device code."""

    def __init__(self, lib, w, h):
        self.lib, self.w, self.h = lib, w, h
        n = w * h
        self.irr, self.prev_irr, self.temp_irr = (np.zeros((n, 4), np.uint32) for _ in range(3))  # 16-byte records
        self.moments, self.prev_moments = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        self.hist, self.hist_prev = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self.cur = None
        self.passes, self.channel = 5, 0

    def set_camera(self, proj, inv, pos, reset=False):
        cur = (np.asarray(proj, np.float32).copy(), np.asarray(inv, np.float32).copy(), np.asarray(pos, np.float64).copy())
        self.history = self.cur if self.cur is not None else cur
        self.cur = cur
        self.moments, self.prev_moments = self.prev_moments, self.moments
        self.reset = reset

    def denoise_present(self, tiles):
        p = lambda a: a.ctypes.data  # noqa: E731
        w, h = self.w, self.h
        t = np.ascontiguousarray(tiles).view(np.uint32).reshape(-1)
        rgba = np.zeros((h, w), np.uint32)
        if self.channel != 4:
            delta = (self.cur[2] - self.history[2]).astype(np.float32)
            self.hist, self.hist_prev = self.hist_prev, self.hist
            self.lib.emu_reproject(p(t), p(self.prev_irr), p(self.prev_moments), p(self.hist_prev), p(self.irr), p(self.moments), p(self.hist),
                                   p(self.cur[1]), p(self.history[0]), p(self.history[1]), p(delta), w, h, int(self.reset))
            if self.passes > 0:
                self.lib.emu_variance(p(self.irr), p(self.hist), p(self.temp_irr), w, h)
                for i in range(self.passes):
                    src = self.prev_irr if i == 1 else (self.temp_irr if i % 2 == 0 else self.irr)
                    dst = self.irr if i % 2 == 0 else self.temp_irr
                    self.lib.emu_atrous(p(src), p(dst), w, h, i)
                    if i == 0:
                        self.prev_irr, self.irr = self.irr, self.prev_irr
                if self.passes % 2 != 0:
                    self.temp_irr, self.irr = self.irr, self.temp_irr
        else:
            self.lib.emu_blit_only(p(t), p(self.irr), p(self.prev_irr), w, h)
        self.lib.emu_present(p(t), p(self.irr), p(rgba), w, h, self.channel)
        if self.passes == 0:
            self.prev_irr, self.irr = self.irr, self.prev_irr
        return rgba


def _half4(records):  # the f16 half of a record buffer as (n, 4) u16, like PostOracle.read
    return np.ascontiguousarray(records[:, :2]).view(np.uint16).reshape(-1, 4)


def _run(emu, seq, w, h, passes, channel_at=None, reset_at=()):
    from oracle import pypostoracle as pp

    gb, orc = EmulatedGBuffer(emu, w, h), pp.PostOracle(w, h)
    gb.passes = passes
    orc.set_passes(passes)
    for f, (proj, inv, pos, tiles) in enumerate(seq):
        ch = (channel_at or {}).get(f, 0)
        gb.channel = ch
        orc.set_debug_channel(ch)
        gb.set_camera(proj, inv, pos, reset=f in reset_at)
        orc.set_camera(proj, inv, pos, reset_history=f in reset_at)
        got, want = gb.denoise_present(tiles), orc.denoise_present(tiles)
        assert np.array_equal(got, want), f"passes={passes} frame {f}: presented image differs at {(got != want).sum()} pixels"
        for name, a, b in (("irr", gb.irr, orc.IRR), ("prev", gb.prev_irr, orc.PREV_IRR), ("temp", gb.temp_irr, orc.TEMP_IRR)):
            assert np.array_equal(_half4(a), orc.read(b)), f"passes={passes} frame {f}: {name}"
        assert np.array_equal(gb.moments.view(np.uint16).reshape(-1, 2), orc.read(orc.MOMENTS)), f"frame {f}: moments"
        assert np.array_equal(gb.hist, orc.read(orc.HIST)), f"frame {f}: history length"
        # the geometry half of the records that carry "PrevIrradiance" is this frame's depth / albedo (what the next frame's
        # Reproject needs as PrevDepthTex / PrevAlbedoTex)
        assert np.array_equal(gb.prev_irr[:, 2].view(np.float32), orc.read(orc.DEPTH)) and np.array_equal(gb.prev_irr[:, 3], orc.read(orc.ALBEDO))


@pytest.mark.parametrize("passes", [0, 1, 2, 3, 4, 5])
def test_kernel_source_equals_oracle_on_synthetic_sequences(emu, passes):
    w, h = 96, 64
    _run(emu, pu.synthetic_sequence(w, h, 5, seed=100 + passes), w, h, passes, reset_at=(3,))


def test_kernel_source_equals_oracle_with_channel_switches_and_large_dilations(emu):
    w, h = 160, 96  # interior CTAs exist for every pass (radius 32 at pass 4)
    seq = pu.synthetic_sequence(w, h, 5, seed=321)
    _run(emu, seq, w, h, 5, channel_at={2: 4, 3: 2})
    _run(emu, seq, w, h, 0, channel_at={1: 4})


def test_kernel_source_equals_oracle_on_traced_frames(emu, hash_oracle, shading_inputs):
    """Frames traced by the traversal ORACLE (1 bounce, blue noise, sky) from a moving camera."""
    from scenes import camera
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    hash_oracle.set_blue_noise(bn)
    hash_oracle.set_sky(desc, tex)
    w, h = 192, 108
    seq = []
    for f in range(4):
        cam = camera.Camera(pos=(96.3 + 0.6 * f, 90.2 + 0.1 * f, 20.7 + 0.4 * f), yaw=0.2 + 0.01 * f, pitch=-0.45)
        proj, inv, wo, frac = cam.matrices(w, h)
        out = hash_oracle.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=f + 1, bounces=1))[0]
        seq.append((proj, inv, cam.pos.copy(), np.frombuffer(out.tobytes(), dtype=capi.TILE_DTYPE).copy()))
    _run(emu, seq, w, h, 5)
