"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libref_cpu.so = the reference's
CpuRenderer.cpp / VoxelMap.cpp compiled from /root/reference, see oracle/ref_harness.cpp).

Run here (the build container has /root/reference):   python tests/golden/make_golden.py
The fixtures travel; tests compare the oracle (CPU, everywhere) and the CUDA path (-m gpu) to them.

Inputs are integer-reproducible: the hash terrain (scenes/terrain.py, pure integer math), seeded
numpy rays, matrices stored in the fixture.  Outputs are what the reference computed:
  ref_trace_lane.npz   RayCast with one active lane per packet (the lane-wise semantics)
  ref_trace_packet.npz RayCast with 16 rays per packet (carries the packet-coupled quirks Q2/Q10)
  ref_hit_query.npz    VoxelMap::RayCast (fp64 picking)
  ref_storage.npz      FlatVoxelStorage after SyncBuffers for a few sectors (masks, cell masks)
  ref_misc.npz         Material::GetEncoded, pixel-format packers, sincos_2pi, blue-noise tiles
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from conftest import random_rays  # noqa: E402
from oracle import refharness  # noqa: E402
from scenes import shading, terrain  # noqa: E402

OUT = Path(__file__).resolve().parent
SCENE_ARGS = dict(nx=6, ny=4, nz=6, seed=77)


def edge_rays(rng, n, wo):
    o, d = random_rays(rng, n, 192, 128, wo)
    specials = np.array([0.0, -0.0, 1e-40, -1e-40, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-30, 3e38], np.float32)
    for k in range(n // 2):
        axis = rng.integers(0, 3)
        d[k, axis] = specials[rng.integers(0, specials.size)]
        if k % 3 == 0:
            d[k, (axis + 1) % 3] = specials[rng.integers(0, specials.size)]
        if k % 7 == 0:
            d[k] = specials[rng.integers(0, specials.size, 3)]
    o[n // 2 : n // 2 + n // 8] = np.round(o[n // 2 : n // 2 + n // 8])
    o[n // 2 + n // 8 : n // 2 + n // 8 + 64] *= np.float32(1e6)
    o[n // 2 + n // 8 + 64 : n // 2 + n // 8 + 96] = np.float32(np.nan)
    return o, d


def main():
    scene = terrain.terrain_hash(**SCENE_ARGS)
    recs = terrain.scene_records(scene)
    ref = refharness.RefMap()
    ref.set_palette(scene["palette"])
    ref.sync(recs)
    digest = terrain.scene_digest(scene)

    rng = np.random.default_rng(20261017)
    wo = np.array([64, 40, 64], np.int32)
    o1, d1 = random_rays(rng, 6144, 192, 128, wo)
    o2, d2 = edge_rays(rng, 2048, wo)
    o, d = np.concatenate([o1, o2]), np.concatenate([d1, d2])
    lane = ref.trace(o, d, wo, lanes_per_packet=1)
    np.savez_compressed(OUT / "ref_trace_lane.npz", origin=o, dir=d, world_origin=wo, hits=lane, scene_digest=digest)
    packet = ref.trace(o, d, wo, lanes_per_packet=16)
    np.savez_compressed(OUT / "ref_trace_packet.npz", origin=o, dir=d, world_origin=wo, hits=packet, scene_digest=digest)

    n = 4096
    qo = np.stack([rng.uniform(0, 192, n), rng.uniform(0, 128, n), rng.uniform(0, 192, n)], 1)
    qd = rng.normal(size=(n, 3))
    qd /= np.linalg.norm(qd, axis=1, keepdims=True)
    np.savez_compressed(OUT / "ref_hit_query.npz", origin=qo, dir=qd, hits=ref.hit_query(qo, qd), scene_digest=digest)

    keys = sorted(scene["sectors"].keys())[::9][:12]
    masks, cells = [], []
    for k in keys:
        m, b, c = ref.read_sector(*k)
        masks.append(m)
        cells.append(c)
    np.savez_compressed(OUT / "ref_storage.npz", keys=np.array(keys, np.int32), masks=np.array(masks, np.uint64), cells=np.stack(cells), scene_digest=digest)

    mats = [(int(r), int(g), int(b), int(f), float(e)) for r, g, b, f, e in zip(rng.integers(0, 256, 64), rng.integers(0, 256, 64), rng.integers(0, 256, 64), rng.integers(0, 256, 64), rng.choice([0.0, 0.5, 0.8, 1.0, 10.0, 123.456, 1e-6, 70000.0], 64))]
    enc = np.array([refharness.encode_material(*m) for m in mats], np.uint64)
    lib = refharness.load()
    cols = rng.random((256, 4)).astype(np.float32) * np.float32(1.6) - np.float32(0.2)
    rgba8 = np.array([lib.ref_pack_rgba8(*[float(v) for v in c]) for c in cols], np.uint32)
    hdr = (rng.random((256, 3)) ** 4 * 300).astype(np.float32)
    r11 = np.array([lib.ref_pack_r11g11b10f(*[float(v) for v in c]) for c in hdr], np.uint32)
    hv = np.concatenate([(rng.random((250, 2)) ** 3 * 100).astype(np.float32), np.array([[0, -0.0], [1e-8, 65504], [65520, 1e5], [np.inf, -np.inf], [6e-8, 6.1e-5], [np.nan, 1]], np.float32)])
    rg16 = np.array([lib.ref_pack_rg16f(float(a), float(b)) for a, b in hv], np.uint32)
    xs = np.concatenate([rng.random(200).astype(np.float32), np.arange(0, 256, dtype=np.float32) * np.float32(1.0 / 255)])
    sc = np.array([refharness.sincos_2pi(x) for x in xs], np.float32)
    bn, _ = shading.load_blue_noise()
    ref.set_blue_noise(bn)
    bn_q = np.array([(int(rng.integers(0, 4000)), int(rng.integers(0, 2200)), int(rng.integers(0, 200)), int(rng.integers(0, 6))) for _ in range(64)], np.uint32)
    bn_tiles = np.stack([refharness.blue_noise_tile(*[int(v) for v in q]) for q in bn_q])
    np.savez_compressed(
        OUT / "ref_misc.npz", materials=np.array(mats, np.float64), materials_encoded=enc, rgba8_in=cols, rgba8=rgba8, r11_in=hdr, r11=r11,
        rg16_in=hv, rg16=rg16, sincos_in=xs, sincos=sc, bn_queries=bn_q, bn_tiles=bn_tiles, bn_table_sha=np.frombuffer(__import__("hashlib").sha256(bn.tobytes()).digest(), np.uint8),
    )
    print("golden written:", [p.name for p in sorted(OUT.glob("*.npz"))], "scene", digest[:16])


if __name__ == "__main__":
    main()
