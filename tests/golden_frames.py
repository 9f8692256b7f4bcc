"""Shared by the CPU and GPU frame-parity tests: the golden G-buffers the REFERENCE rendered (tests/golden/ref_frames.npz, written by
tests/golden/make_golden_frames.py from oracle/_ref = the reference's own RenderRow)."""
from __future__ import annotations

import hashlib
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden" / "ref_frames.npz"
PLANES = ("albedo", "depth", "irr_rg", "irr_bx")
HASH_NAMES = ("hash_b0", "hash_b1", "hash_b2", "hash_b3", "hash_inside_solid", "hash_outside_view")
BENCH_NAMES = ("bench_720p_b0", "bench_720p_b1", "bench_4k_b0")


def load():
    return np.load(GOLDEN)


def assets_match(z):
    """The blue-noise table and sky cube the golden frames were rendered with (scenes/_ref, converted from the reference's assets)."""
    from scenes import shading

    bn = shading.load_blue_noise()[0]
    _, texels, _ = shading.load_sky()
    return (hashlib.sha256(bn.tobytes()).hexdigest() == str(z["bn_sha"]) and
            hashlib.sha256(np.ascontiguousarray(texels).tobytes()).hexdigest() == str(z["sky_sha"]))


def frame_of(z, name):
    from voxelrt_b200 import capi

    w, h, bounces, frame_no, wx, wy, wz = (int(v) for v in z[name + "_frame"])
    m = z[name + "_mats"]
    return capi.make_frame(w, h, m[16:32], m[0:16], (wx, wy, wz), m[32:35], frame_no=frame_no, bounces=bounces)


def assert_tiles_equal(got, want, what):
    """Byte equality of two tile framebuffers, plane by plane, with the differing pixels counted in the message."""
    for k in PLANES:
        a, b = np.ascontiguousarray(got[k]).view(np.uint32), np.ascontiguousarray(want[k]).view(np.uint32)
        bad = np.nonzero(a.ravel() != b.ravel())[0]
        assert bad.size == 0, f"{what}: plane {k} differs in {bad.size} of {a.size} pixels, first tile-order indices {bad[:6]}"


def assert_digests_equal(got, z, name):
    for i, k in enumerate(PLANES):
        d = hashlib.sha256(np.ascontiguousarray(got[k]).tobytes()).hexdigest()
        assert d == str(z[name + "_sha"][i]), f"{name}: SHA-256 of plane {k} differs from the reference-rendered frame"
    assert int((got["depth"] >= 0).sum()) == int(z[name + "_hits"])
