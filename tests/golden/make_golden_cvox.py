"""Writes tests/golden/ref_cvox_small.dat — a "cvox 0004" file produced by the REFERENCE's own VoxelMap::Serialize
(oracle/_ref) — and ref_cvox_small.npz, the content that went in.  Run where /root/reference exists:
    python tests/golden/make_golden_cvox.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent


def main():
    from oracle import refharness
    from scenes import terrain

    rng = np.random.default_rng(12)
    sectors = {}
    for key in [(0, 0, 0), (3, 1, 2), (-5, -1, 7), (2047, 127, -2048)]:
        mask = int(rng.integers(1, 1 << 40)) | 1
        k = bin(mask).count("1")
        b = rng.integers(0, 256, (k, 512), dtype=np.uint8)
        b[rng.random((k, 512)) < 0.7] = 0
        sectors[key] = (mask, b)
    mats = [(int(rng.integers(0, 256)), int(rng.integers(0, 256)), int(rng.integers(0, 256)), int(rng.integers(0, 256)), float(np.float32(rng.random() * 4))) for _ in range(256)]
    ref = refharness.RefMap()
    ref.sync(terrain.scene_records({"sectors": sectors}))
    ref.set_materials(mats)
    ref.serialize(OUT / "ref_cvox_small.dat")
    keys = sorted(sectors)
    np.savez_compressed(OUT / "ref_cvox_small.npz", keys=np.array(keys, np.int32), masks=np.array([sectors[k][0] for k in keys], np.uint64),
                        bricks=np.concatenate([sectors[k][1] for k in keys]), materials=np.array(mats, np.float64))
    print("written", (OUT / "ref_cvox_small.dat").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
