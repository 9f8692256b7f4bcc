"""Golden G-BUFFERS rendered by the REFERENCE ITSELF: RenderRow of the reference's CpuRenderer.cpp (oracle/_ref, 16-lane AVX-512 packets,
rsqrt14 / rcp14) over whole frames.  Run in the build container:   python tests/golden/make_golden_frames.py

  ref_frames.npz   hash terrain (integer-reproducible scene): full tile framebuffers for 0 / 1 / 2 / 3 bounces, a camera inside solid
                   rock and one outside the view (packet quirks Q2 / Q10 and the never-masked tail of RenderRow all occur);
                   bench terrain (BASELINE configs[0] / [1] scene and camera): SHA-256 of every plane of the 1280x720 frame
                   (0 and 1 bounce) and of the 3840x2160 primary frame — the frames themselves are 15 / 133 MB.
Matrices, frame numbers and the hashes of the blue-noise table and sky cube used are stored with the frames.  The oracle (CPU) and the
CUDA path (-m gpu) must reproduce every byte.
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import refharness  # noqa: E402
from scenes import camera, shading, terrain  # noqa: E402
from voxelrt_b200 import capi  # noqa: E402

OUT = Path(__file__).resolve().parent
PLANES = ("albedo", "depth", "irr_rg", "irr_bx")
HASH_FRAMES = [
    # name, camera, w, h, bounces, frame_no
    ("hash_b0", dict(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45), 256, 144, 0, 1),
    ("hash_b1", dict(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45), 256, 144, 1, 1),
    ("hash_b2", dict(pos=(96.3, 90.2, 20.7), yaw=0.2, pitch=-0.45), 256, 144, 2, 7),
    ("hash_b3", dict(pos=(30.7, 70.1, 150.2), yaw=2.4, pitch=-0.3), 132, 76, 3, 70),
    ("hash_inside_solid", dict(pos=(60.3, 20.2, 40.7), yaw=1.2, pitch=-0.3), 128, 72, 2, 3),
    ("hash_outside_view", dict(pos=(-50.3, 90.2, 20.7), yaw=1.2, pitch=-0.3), 128, 72, 1, 3),
]
BENCH_FRAMES = [("bench_720p_b0", 1280, 720, 0, 1), ("bench_720p_b1", 1280, 720, 1, 1), ("bench_4k_b0", 3840, 2160, 0, 1)]


def assets():
    bn = shading.load_blue_noise()[0]
    desc, texels, _ = shading.load_sky()
    return bn, desc, texels, hashlib.sha256(bn.tobytes()).hexdigest(), hashlib.sha256(np.ascontiguousarray(texels).tobytes()).hexdigest()


def plane_digests(tiles):
    return {k: hashlib.sha256(np.ascontiguousarray(tiles[k]).tobytes()).hexdigest() for k in PLANES}


def main():
    bn, desc, texels, bn_sha, sky_sha = assets()
    out = {"bn_sha": bn_sha, "sky_sha": sky_sha}

    def ref_for(scene):
        ref = refharness.RefMap()
        ref.set_palette(scene["palette"])
        ref.sync(terrain.scene_records(scene))
        ref.set_blue_noise(bn)
        ref.set_sky(desc, texels)
        return ref

    scene = terrain.terrain_hash(6, 4, 6, seed=77)
    out["hash_scene_digest"] = terrain.scene_digest(scene)
    ref = ref_for(scene)
    for name, cam_kw, w, h, bounces, frame_no in HASH_FRAMES:
        proj, inv, wo, frac = camera.Camera(**cam_kw).matrices(w, h)
        tiles, _ = ref.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=frame_no, bounces=bounces))
        out[name + "_tiles"] = tiles.copy()
        out[name + "_frame"] = np.concatenate([[w, h, bounces, frame_no], wo]).astype(np.int64)
        out[name + "_mats"] = np.concatenate([proj, inv, frac]).astype(np.float32)
    ref.close()

    scene = terrain.bench_terrain()
    out["bench_scene_digest"] = terrain.scene_digest(scene)
    ref = ref_for(scene)
    for name, w, h, bounces, frame_no in BENCH_FRAMES:
        proj, inv, wo, frac = camera.Camera().matrices(w, h)
        tiles, _ = ref.render(capi.make_frame(w, h, inv, proj, wo, frac, frame_no=frame_no, bounces=bounces))
        d = plane_digests(tiles)
        out[name + "_sha"] = np.array([d[k] for k in PLANES])
        out[name + "_hits"] = np.int64((tiles["depth"] >= 0).sum())
        out[name + "_frame"] = np.concatenate([[w, h, bounces, frame_no], wo]).astype(np.int64)
        out[name + "_mats"] = np.concatenate([proj, inv, frac]).astype(np.float32)
    np.savez_compressed(OUT / "ref_frames.npz", **out)
    print("wrote ref_frames.npz", (OUT / "ref_frames.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
