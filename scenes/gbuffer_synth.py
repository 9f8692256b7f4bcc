"""Synthetic inputs of the GBuffer (denoise + present) step: tile framebuffers and frame sequences (tests, bench)."""
from __future__ import annotations

import numpy as np

from voxelrt_b200.capi import TILE_DTYPE


def make_tiles(albedo_rgb, normal_code, depth, irr):
    """Pack per-pixel planes into the reference's Framebuffer::Tile layout (4x4 tiles, CpuRenderer.cpp:299-309).
    albedo_rgb u8 [h,w,3]; normal_code u8 [h,w] = (nx+1)|(ny+1)<<2|(nz+1)<<4; depth f32 [h,w]; irr f32 [h,w,3]."""
    h, w = depth.shape
    assert w % 4 == 0 and h % 4 == 0
    a = (
        albedo_rgb[..., 0].astype(np.uint32)
        | albedo_rgb[..., 1].astype(np.uint32) << 8
        | albedo_rgb[..., 2].astype(np.uint32) << 16
        | normal_code.astype(np.uint32) << 24
    )
    hf = irr.astype(np.float32).astype(np.float16).view(np.uint16).astype(np.uint32)
    rg = hf[..., 0] | hf[..., 1] << 16
    bx = hf[..., 2]

    def tile(plane):  # [h,w] -> [h/4, w/4, 16] with lane = (x&3) | (y&3)<<2
        return plane.reshape(h // 4, 4, w // 4, 4).transpose(0, 2, 1, 3).reshape(h // 4, w // 4, 16)

    t = np.zeros((h // 4, w // 4), TILE_DTYPE)
    t["albedo"] = tile(a)
    t["depth"] = tile(depth.astype(np.float32))
    t["irr_rg"] = tile(rg)
    t["irr_bx"] = tile(bx)
    return t.reshape(-1)


def untile(tiles, w, h, field):
    t = tiles.view(TILE_DTYPE).reshape(h // 4, w // 4)[field]
    return t.reshape(h // 4, w // 4, 4, 4).transpose(0, 2, 1, 3).reshape(h, w)


def normal_code(nx, ny, nz):
    return (nx + 1) | (ny + 1) << 2 | (nz + 1) << 4


def f16_to_f32(bits):
    return np.asarray(bits, np.uint16).view(np.float16).astype(np.float32)


def synthetic_sequence(w, h, frames, seed, moving=True, sky_fraction=0.15):
    """A seeded sequence of (camera, tiles) with geometry that is consistent from frame to frame: the view looks
    at axis-aligned slabs (per-column depth steps and normal changes), the camera pans a little every frame, and the
    irradiance is a smooth signal plus per-frame noise.  Returns [(proj, inv_proj, position, tiles)]."""
    from scenes import camera

    rng = np.random.default_rng(seed)
    out = []
    # per-column geometry classes, fixed over the sequence
    cls = (np.arange(w) * 7 // w) % 3
    ncode = np.array([normal_code(0, 1, 0), normal_code(-1, 0, 0), normal_code(0, 0, 1)], np.uint8)[cls]
    base_depth = np.array([0.9990, 0.9995, 0.9985], np.float32)[cls]
    sky_rows = int(h * sky_fraction)
    albedo = rng.integers(30, 255, size=(h, w, 3), dtype=np.uint8)
    albedo = np.repeat(np.repeat(albedo[::4, ::4], 4, 0), 4, 1)[:h, :w]
    smooth = 0.6 + 0.4 * np.sin(np.arange(w)[None, :] * 0.11) * np.cos(np.arange(h)[:, None] * 0.07)
    for f in range(frames):
        cam = camera.Camera(pos=(100.25 + (0.37 * f if moving else 0.0), 80.5, 64.75 - (0.21 * f if moving else 0.0)),
                            yaw=0.4 + (0.004 * f if moving else 0.0), pitch=-0.3)
        proj, inv, _, _ = cam.matrices(w, h)
        depth = np.broadcast_to(base_depth[None, :], (h, w)).copy()
        depth += (np.arange(h, dtype=np.float32)[:, None] * np.float32(1e-6))
        depth[:sky_rows] = -1.0
        irr = (smooth[..., None] * np.array([1.0, 0.8, 0.6])) * (1.0 + 0.5 * rng.standard_normal((h, w, 3)))
        irr = np.clip(irr, 0.0, 8.0).astype(np.float32)
        nc = np.broadcast_to(ncode[None, :], (h, w)).copy()
        nc[:sky_rows] = normal_code(0, 0, 0)
        tiles = make_tiles(albedo, nc, depth, irr)
        out.append((proj, inv, cam.pos.copy(), tiles))
    return out
