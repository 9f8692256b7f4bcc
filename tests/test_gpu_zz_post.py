"""GPU parity of the GBuffer step (include/voxelrt_b200_post.h: blit + reprojection + SVGF + present) against the
image-space oracle, through the C ABI.  The bar is BIT-EXACT for every plane and the presented RGBA8 image: both sides
evaluate the same canonical fp32 operation order with the same polynomial exp/log (see oracle/vrt_post_oracle.c), and f16
stores round to nearest even on both.  (The oracle restates the reference's GLSL; that restatement is unpinned.)"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from scenes import gbuffer_synth as pu

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _compare_planes(gb, orc, what):
    from voxelrt_b200 import post

    for plane_g, plane_o, name in (
        (post.VRT_PLANE_IRRADIANCE, orc.IRR, "IrradianceTex"),
        (post.VRT_PLANE_PREV_IRRADIANCE, orc.PREV_IRR, "PrevIrradianceTex"),
        (post.VRT_PLANE_TEMP_IRRADIANCE, orc.TEMP_IRR, "TempIrradianceTex"),
    ):
        g = gb.read(plane_g)["irr"]
        o = orc.read(plane_o)
        bad = np.nonzero((g != o).any(axis=1))[0]
        assert bad.size == 0, f"{what}: {name} differs at {bad.size} pixels, first {bad[:5]}: {g[bad[:3]]} vs {o[bad[:3]]}"
    assert np.array_equal(gb.read(post.VRT_PLANE_MOMENTS), orc.read(orc.MOMENTS)), f"{what}: MomentsTex"
    assert np.array_equal(gb.read(post.VRT_PLANE_HISTORY_LEN), orc.read(orc.HIST)), f"{what}: HistoryLenTex"


def _run_both(seq, w, h, passes, channel=0, reset_at=(), channel_at=None, check_planes=True, what=""):
    from oracle import pypostoracle as pp
    from voxelrt_b200 import post

    gb = post.GBuffer(0)
    orc = pp.PostOracle(w, h)
    gb.set_passes(passes)
    orc.set_passes(passes)
    for f, (proj, inv, pos, tiles) in enumerate(seq):
        ch = channel_at.get(f, channel) if channel_at else channel
        gb.set_debug_channel(ch)
        orc.set_debug_channel(ch)
        gb.set_camera(post.make_camera(w, h, proj, inv, pos, reset_history=f in reset_at))
        orc.set_camera(proj, inv, pos, reset_history=f in reset_at)
        img_g = gb.denoise_present(tiles)
        img_o = orc.denoise_present(tiles)
        bad = np.argwhere(img_g != img_o)
        assert bad.size == 0, f"{what} frame {f}: presented image differs at {len(bad)} pixels, first {bad[:4].tolist()}"
        want_launches = 2 if ch == 4 else 2 + (1 + passes if passes else 0)
        assert gb.last_launches() == want_launches
        if check_planes:
            _compare_planes(gb, orc, f"{what} frame {f}")
    gb.close()


@pytest.mark.parametrize("passes", [0, 1, 2, 3, 4, 5])
def test_synthetic_sequence_bit_exact(passes):
    w, h = 96, 64
    seq = pu.synthetic_sequence(w, h, 6, seed=100 + passes)
    _run_both(seq, w, h, passes, reset_at=(4,), what=f"passes={passes}")


def test_full_size_3840x2160_bit_exact():
    """BASELINE's frame size: two 4K frames (the second one reprojects), five passes, every plane and the image bit-exact —
    covers interior CTAs of every pass and 64-bit indexing.  The oracle runs its passes on all host threads."""
    w, h = 3840, 2160
    seq = pu.synthetic_sequence(w, h, 2, seed=77)
    _run_both(seq, w, h, 5, what="4K")


def test_static_camera_long_history_bit_exact():
    w, h = 64, 48
    seq = pu.synthetic_sequence(w, h, 70, seed=7, moving=False)
    _run_both(seq, w, h, 5, check_planes=False, what="static")
    _run_both(seq[:12], w, h, 2, what="static N=2")


@pytest.mark.parametrize("channel", [1, 2, 3, 4, 5])
def test_debug_channels_bit_exact(channel):
    w, h = 64, 48
    seq = pu.synthetic_sequence(w, h, 3, seed=300 + channel)
    _run_both(seq, w, h, 5, channel=channel, what=f"channel={channel}")


def test_switching_channel_and_passes_between_frames():
    """TraversalIters frames skip the denoiser but still advance the 'previous frame' geometry (GBuffer.h:52-56,90)."""
    w, h = 64, 48
    seq = pu.synthetic_sequence(w, h, 8, seed=400)
    _run_both(seq, w, h, 5, channel_at={2: 4, 3: 4, 5: 2}, what="channel switches")
    _run_both(seq, w, h, 0, channel_at={1: 4, 4: 4}, what="channel switches N=0")


def test_regression_fixture_on_gpu():
    from voxelrt_b200 import post

    z = np.load(GOLDEN / "post_sequence.npz")
    w, h, frames, passes = (int(z[k]) for k in ("w", "h", "frames", "passes"))
    seq = pu.synthetic_sequence(w, h, frames, seed=int(z["seed"]))
    gb = post.GBuffer(0)
    gb.set_passes(passes)
    for f, (proj, inv, pos, tiles) in enumerate(seq):
        gb.set_camera(post.make_camera(w, h, proj, inv, pos, reset_history=(f == 3)))
        img = gb.denoise_present(tiles)
        assert np.array_equal(img, z["rgba"][f]), f"frame {f}"
    assert np.array_equal(gb.read(post.VRT_PLANE_HISTORY_LEN), z["hist"])
    assert np.array_equal(gb.read(post.VRT_PLANE_MOMENTS), z["moments"])
    assert np.array_equal(gb.read(post.VRT_PLANE_PREV_IRRADIANCE)["irr"], z["prev_irr"])


def test_traced_frames_bit_exact(hash_scene, shading_inputs):
    """The real thing: frames traced by the CUDA path (1 bounce, blue-noise, sky) from a moving camera, denoised on the GPU
    and by the oracle from the same tile bytes."""
    from conftest import ctx_for
    from scenes import camera
    from voxelrt_b200 import capi

    (bn, _), (desc, tex, _) = shading_inputs
    ctx = ctx_for(hash_scene)
    ctx.set_blue_noise(bn)
    ctx.set_sky(desc, tex)
    w, h = 256, 144
    seq = []
    for f in range(5):
        cam = camera.Camera(pos=(96.3 + 0.6 * f, 90.2 + 0.1 * f, 20.7 + 0.4 * f), yaw=0.2 + 0.01 * f, pitch=-0.45)
        proj, inv, wo, frac = cam.matrices(w, h)
        out, _ = ctx.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=f + 1, bounces=1))
        seq.append((proj, inv, cam.pos.copy(), np.frombuffer(out.tobytes(), dtype=capi.TILE_DTYPE).copy()))
    _run_both(seq, w, h, 5, what="traced")
    _run_both(seq, w, h, 0, what="traced N=0")


def test_device_pointer_path_matches_host_path():
    import torch

    from voxelrt_b200 import post

    w, h = 128, 64
    seq = pu.synthetic_sequence(w, h, 4, seed=500)
    gb_h, gb_d = post.GBuffer(0), post.GBuffer(0)
    stream = torch.cuda.Stream()
    for proj, inv, pos, tiles in seq:
        cam = post.make_camera(w, h, proj, inv, pos)
        gb_h.set_camera(cam)
        want = gb_h.denoise_present(tiles)
        gb_d.set_camera(cam)
        d_tiles = torch.from_numpy(tiles.view(np.int32).copy()).cuda()
        d_out = torch.zeros(w * h, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        gb_d.denoise_present_device(d_tiles.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        assert np.array_equal(d_out.cpu().numpy().view(np.uint32).reshape(h, w), want)


def test_size_change_drops_history_and_errors_are_loud():
    from voxelrt_b200 import capi, post

    gb = post.GBuffer(0)
    w, h = 64, 48
    seq = pu.synthetic_sequence(w, h, 3, seed=600, moving=False)
    with pytest.raises(capi.VrtError):  # no camera yet
        gb.lib.vrt_gbuffer_set_passes(gb.h, 5)
        gb.width, gb.height = w, h
        gb.denoise_present(seq[0][3])
    for proj, inv, pos, tiles in seq:
        gb.set_camera(post.make_camera(w, h, proj, inv, pos))
        gb.denoise_present(tiles)
    assert gb.read(post.VRT_PLANE_HISTORY_LEN).max() == 2
    w2, h2 = 96, 64
    proj, inv, pos, tiles = pu.synthetic_sequence(w2, h2, 1, seed=601)[0]
    gb.set_camera(post.make_camera(w2, h2, proj, inv, pos))
    gb.denoise_present(tiles)
    assert gb.read(post.VRT_PLANE_HISTORY_LEN).max() == 0
    with pytest.raises(capi.VrtError):
        gb.set_passes(6)
    with pytest.raises(capi.VrtError):
        gb.set_camera(post.make_camera(30, 20, proj, inv, pos))


def test_native_adapter_denoise_and_present():
    """vrt_host::B200Renderer with DenoiseAndPresent = true (C++ host, tests/native/test_host.cpp --gpu-present): four
    frames from a moving camera, presented image == traversal oracle + image-space oracle chained."""
    import subprocess

    exe = Path(__file__).resolve().parent / "native" / "test_host"
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(exe.parent)], check=True)
    r = subprocess.run([str(exe), "--gpu-present"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "native tests: ok" in r.stdout and "gpu-present:" in r.stdout, r.stdout + r.stderr
