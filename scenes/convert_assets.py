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


def read_hdr(path):
    """Radiance RGBE panorama -> float32[h, w, 3] (what stbi_loadf returns, ImageHelpers.cpp:60-62)."""
    import ctypes as C

    raw = path.read_bytes()
    end = raw.index(b"\n\n") + 2
    nl = raw.index(b"\n", end)
    res = raw[end:nl].split()
    assert res[0] == b"-Y" and res[2] == b"+X", res
    h, w = int(res[1]), int(res[3])
    body = np.frombuffer(raw, np.uint8, offset=nl + 1)
    out = np.zeros((h, w, 4), np.uint8)
    lib = C.CDLL(str(HERE / "_ref" / "libvoxelizer.so"))
    lib.rgbe_decode.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.rgbe_decode.restype = C.c_int
    if lib.rgbe_decode(body.ctypes.data, body.size, w, h, out.ctypes.data) != 0:
        raise ValueError(f"{path}: not a new-style RLE .hdr")
    e = out[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0.0)).astype(np.float32)
    return out[..., :3].astype(np.float32) * scale[..., None]


def convert_sky(face_mips=8) -> bool:
    """The reference's sky: assets/skyboxes/evening_road_01_puresky_4k.hdr -> R11G11B10f cube, face = width / 4, box-filtered
    mips (LoadCubemapFromPanoramaHDR + GenerateMips, SwRast/ImageHelpers.cpp:87-122, Texture.h:683-704) -> scenes/_ref/sky_<face>.bin
    (compact: the used texels of the six layers; scenes/shading.load_sky pads them to the Texture2D layer stride).
    Float bilinear filtering stands in for the reference's fixed-point lerp: an INPUT table, identical on both sides of every test."""
    from scenes import shading

    src = REF / "assets/skyboxes/evening_road_01_puresky_4k.hdr"
    if not src.exists() or not (HERE / "_ref" / "libvoxelizer.so").exists():
        return False
    pano = read_hdr(src)
    face = pano.shape[1] // 4
    dst = HERE / "_ref" / f"sky_{face}.bin"
    if dst.exists() and dst.stat().st_mtime >= src.stat().st_mtime:
        return True
    pano = shading.unpack_r11g11b10f(shading.pack_r11g11b10f(pano))  # the panorama itself is stored packed (LoadImageHDR)
    ph, pw = pano.shape[:2]
    desc = shading.sky_layout(face, face_mips)
    layer_size = desc.mip_offset[desc.mip_levels - 1] + (face >> (desc.mip_levels - 1)) ** 2
    layer_size = (layer_size + 15) & ~15
    out = np.zeros((6, layer_size), np.uint32)
    ax = np.arange(face, dtype=np.float32) / np.float32(face - 1)
    u, v = np.meshgrid(ax, ax, indexing="xy")
    for layer in range(6):
        d = shading._unproject(layer, u.astype(np.float64), v.astype(np.float64))
        pu = np.arctan2(d[..., 2], d[..., 0]) / (2 * np.pi) + 0.5
        pv = np.arcsin(-d[..., 1]) / np.pi + 0.5
        fx, fy = pu * pw - 0.5, pv * ph - 0.5
        x0, y0 = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
        wx, wy = (fx - x0)[..., None].astype(np.float32), (fy - y0)[..., None].astype(np.float32)
        x1, y1 = (x0 + 1) % pw, np.clip(y0 + 1, 0, ph - 1)
        x0, y0 = x0 % pw, np.clip(y0, 0, ph - 1)
        img = (pano[y0, x0] * (1 - wx) + pano[y0, x1] * wx) * (1 - wy) + (pano[y1, x0] * (1 - wx) + pano[y1, x1] * wx) * wy
        packed = shading.pack_r11g11b10f(img.astype(np.float32))
        for lvl in range(desc.mip_levels):
            s = face >> lvl
            out[layer, desc.mip_offset[lvl] : desc.mip_offset[lvl] + s * s] = packed.reshape(-1)
            if lvl + 1 < desc.mip_levels:
                t = shading.unpack_r11g11b10f(packed).reshape(s // 2, 2, s // 2, 2, 3)
                avg = (t[:, 0, :, 0] + t[:, 0, :, 1] + t[:, 1, :, 0] + t[:, 1, :, 1]) * np.float32(0.25)
                packed = shading.pack_r11g11b10f(avg)
    dst.parent.mkdir(parents=True, exist_ok=True)
    out.tofile(dst)
    return True


if __name__ == "__main__":
    sys.path.insert(0, str(HERE.parent))
    ok = convert_blue_noise()
    print("blue noise:", "converted" if ok else "reference asset absent")
    ok = convert_sky()
    print("sky cube:", "converted" if ok else "reference asset / decoder absent")
    sys.exit(0)
