"""Converts the reference's shading assets into raw tables under scenes/_ref/ (git-ignored, travels to
the GPU box).  Run by __graft_entry__.build() where /root/reference exists; INPUT data only.

  assets/bluenoise/stbn_vec2_2Dx1D_128x128x64_combined.png  ->  scenes/_ref/bluenoise_rg.bin
      128 x 8192 RGBA PNG; the renderer keeps (R,G) of every texel (VoxelRT/CpuRenderer.cpp:239-248).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")


def convert_blue_noise() -> bool:
    src = REF / "assets/bluenoise/stbn_vec2_2Dx1D_128x128x64_combined.png"
    dst = HERE / "_ref" / "bluenoise_rg.bin"
    if not src.exists():
        return False
    if dst.exists() and dst.stat().st_mtime >= src.stat().st_mtime:
        return True
    from PIL import Image

    img = np.asarray(Image.open(src).convert("RGBA"))
    assert img.shape == (8192, 128, 4), img.shape
    dst.parent.mkdir(parents=True, exist_ok=True)
    np.ascontiguousarray(img[:, :, :2]).tofile(dst)
    return True


if __name__ == "__main__":
    ok = convert_blue_noise()
    print("blue noise:", "converted" if ok else "reference asset absent")
    sys.exit(0)
