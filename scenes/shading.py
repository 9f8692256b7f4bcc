"""Shading INPUT tables (not the product, not the oracle): the blue-noise sequence and the sky cube.

Blue noise: the reference ships `assets/bluenoise/stbn_vec2_2Dx1D_128x128x64_combined.png`
(128 x 8192 RGBA, R and G used, VoxelRT/CpuRenderer.cpp:239-248).  `scenes/convert_assets.py`
(run by __graft_entry__.build() where /root/reference exists) stores its (R,G) bytes as
`scenes/_ref/bluenoise_rg.bin`, which travels to the GPU box; without it a seeded synthetic table of
the same shape is used and the scene name says so.

Sky: swr::HdrTexture2D layout (LibGlimpsw/SwRast/Texture.h:401-446): 6 layers of R11G11B10f texels,
mip chain per layer, layer stride rounded to a power of two.  `scenes/_ref/sky_<face>.bin` holds the
cube built from the reference's panorama when available; otherwise a procedural gradient sky.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from voxelrt_b200.capi import VRT_BLUE_NOISE_BYTES, VrtSkyDesc

HERE = Path(__file__).resolve().parent
BN_FILE = HERE / "_ref" / "bluenoise_rg.bin"


def load_blue_noise():
    """-> (uint8[8192,128,2], source string)."""
    if BN_FILE.exists() and BN_FILE.stat().st_size == VRT_BLUE_NOISE_BYTES:
        return np.fromfile(BN_FILE, np.uint8).reshape(8192, 128, 2), "reference STBN vec2 128x128x64"
    rng = np.random.default_rng(20261017)
    return rng.integers(0, 256, (8192, 128, 2), dtype=np.uint8), "synthetic uniform bytes (reference PNG not converted)"


def sky_layout(face_size: int, mip_levels: int = 8):
    """Texture2D ctor arithmetic, Texture.h:409-433 (VectorWidth = 16)."""
    desc = VrtSkyDesc()
    desc.face_size = face_size
    levels, layer_size = 0, 0
    while levels < min(mip_levels, 16):
        if (face_size >> levels) < 4:
            break
        desc.mip_offset[levels] = layer_size
        layer_size += (face_size >> levels) ** 2
        layer_size = (layer_size + 15) & ~15
        levels += 1
    desc.mip_levels = levels
    desc.layer_shift = int(layer_size).bit_length()
    desc.texel_count = (1 << desc.layer_shift) * 6 + 16
    return desc


def pack_r11g11b10f(rgb: np.ndarray) -> np.ndarray:
    """R11G11B10f::Pack, Texture.h:158-177 (clamp, then truncate the f32 mantissa)."""
    rgb = np.asarray(rgb, np.float32)
    r = np.clip(rgb[..., 0], np.float32(1.0 / (1 << 15)), np.float32(130048.0)).view(np.int32)
    g = np.clip(rgb[..., 1], np.float32(1.0 / (1 << 15)), np.float32(130048.0)).view(np.int32)
    b = np.clip(rgb[..., 2], np.float32(1.0 / (1 << 15)), np.float32(129024.0)).view(np.int32)
    pr = (((r >> 17) & 0x3FFF) - 0x1C00).astype(np.uint32)
    pg = (((g >> 17) & 0x3FFF) - 0x1C00).astype(np.uint32)
    pb = (((b >> 18) & 0x1FFF) - 0x0E00).astype(np.uint32)
    return (pr << np.uint32(21)) | (pg << np.uint32(10)) | pb


def unpack_r11g11b10f(t: np.ndarray) -> np.ndarray:
    t = np.asarray(t, np.uint32)
    r = ((((t >> np.uint32(21)) << np.uint32(17)) & np.uint32(0x0FFE0000)) + np.uint32(0x38000000)).view(np.float32)
    g = ((((t >> np.uint32(10)) << np.uint32(17)) & np.uint32(0x0FFE0000)) + np.uint32(0x38000000)).view(np.float32)
    b = (((t << np.uint32(18)) & np.uint32(0x0FFC0000)) + np.uint32(0x38000000)).view(np.float32)
    return np.stack([r, g, b], axis=-1)


def _unproject(face, u, v):
    """texutil::UnprojectCubemap (Texture.h:291-311): face = axis*2 + negative."""
    w = -1.0 if (face & 1) else 1.0
    axis = face >> 1
    u = u * 2 - 1
    v = v * 2 - 1
    if axis == 0:
        d = np.stack([np.full_like(u, w), v, u], -1)
    elif axis == 1:
        d = np.stack([u, np.full_like(u, w), v], -1)
    else:
        d = np.stack([u, v, np.full_like(u, w)], -1)
    return d / np.linalg.norm(d, axis=-1, keepdims=True)


def procedural_sky(face_size=64):
    """Gradient sky + a sun disc, box-filtered mips.  -> (VrtSkyDesc, uint32 texels)."""
    desc = sky_layout(face_size)
    tex = np.zeros(desc.texel_count, np.uint32)
    for face in range(6):
        ax = (np.arange(face_size, dtype=np.float64)) / (face_size - 1)
        u, v = np.meshgrid(ax, ax, indexing="xy")
        d = _unproject(face, u, v)
        up = np.clip(d[..., 1] * 0.5 + 0.5, 0, 1)
        base = np.stack([0.25 + 0.35 * (1 - up), 0.35 + 0.4 * (1 - up) * up + 0.1, 0.55 + 0.45 * up], -1)
        sun = np.clip((d @ np.array([0.45, 0.6, 0.66])) - 0.985, 0, 1) * 4000.0
        img = (base + sun[..., None]).astype(np.float32)
        for lvl in range(desc.mip_levels):
            s = face_size >> lvl
            off = (face << desc.layer_shift) + desc.mip_offset[lvl]
            tex[off : off + s * s] = pack_r11g11b10f(img).reshape(-1)
            if s >= 8:
                img = img.reshape(s // 2, 2, s // 2, 2, 3).mean(axis=(1, 3)).astype(np.float32)
    return desc, tex


def load_sky(face_size=None):
    """-> (VrtSkyDesc, texels uint32, source string)."""
    cands = sorted((HERE / "_ref").glob("sky_*.bin")) if (HERE / "_ref").is_dir() else []
    for p in cands:
        fs = int(p.stem.split("_")[1])
        if face_size is not None and fs != face_size:
            continue
        desc = sky_layout(fs)
        tex = np.fromfile(p, np.uint32)
        if tex.size == desc.texel_count:
            return desc, tex, f"reference panorama cube {fs}^2 ({p.name})"
        if tex.size % 6 == 0 and tex.size // 6 <= (1 << desc.layer_shift):  # compact file: used texels of each layer
            full = np.zeros(desc.texel_count, np.uint32)
            per = tex.size // 6
            for layer in range(6):
                full[layer << desc.layer_shift : (layer << desc.layer_shift) + per] = tex[layer * per : (layer + 1) * per]
            return desc, full, f"reference panorama cube {fs}^2 ({p.name}: evening_road_01_puresky_4k.hdr)"
    desc, tex = procedural_sky(face_size or 64)
    return desc, tex, "procedural gradient sky (reference HDR not converted)"
