"""An INDEPENDENT restatement of the GLSL renderer's ray casts (Shaders/VoxelTraversal.glsl:92-243) in plain Python over a DENSE
voxel array — occupancy masks are recomputed from the voxels on the fly, nothing is shared with oracle/vrt_oracle.c.  All
arithmetic is numpy float32 scalar arithmetic (IEEE RN, one operation at a time), so tests/test_glsl_oracle.py can ask
orc_trace_glsl for the SAME BITS.  Small cases only (pure-Python loops)."""
from __future__ import annotations

import numpy as np

F = np.float32


class DenseWorld:
    def __init__(self, scene, xz_log2=6, y_log2=4):
        keys = list(scene["sectors"].keys())
        self.nx = max(k[0] for k in keys) + 1
        self.ny = max(k[1] for k in keys) + 1
        self.nz = max(k[2] for k in keys) + 1
        self.vox = np.zeros((self.ny * 32, self.nz * 32, self.nx * 32), np.uint8)  # [y, z, x]
        self.alloc = {}
        for (sx, sy, sz), (mask, bricks) in scene["sectors"].items():
            self.alloc[(sx, sy, sz)] = int(mask)
            k = 0
            for b in range(64):
                if (int(mask) >> b) & 1:
                    bx, bz, by = b & 3, (b >> 2) & 3, b >> 4
                    blk = np.asarray(bricks[k]).reshape(8, 8, 8)  # voxel index x | z<<3 | y<<6 -> [y, z, x]
                    x0, y0, z0 = sx * 32 + bx * 8, sy * 32 + by * 8, sz * 32 + bz * 8
                    self.vox[y0 : y0 + 8, z0 : z0 + 8, x0 : x0 + 8] = blk
                    k += 1
        self.grid = (32 << xz_log2, 32 << y_log2, 32 << xz_log2)
        self.palette = np.asarray(scene["palette"], np.uint64)

    def voxel(self, x, y, z):
        if 0 <= x < self.vox.shape[2] and 0 <= y < self.vox.shape[0] and 0 <= z < self.vox.shape[1]:
            return int(self.vox[y, z, x])
        return 0

    def brick_allocated(self, bx, by, bz):
        m = self.alloc.get((bx >> 2, by >> 2, bz >> 2), 0)
        return (m >> ((bx & 3) | (bz & 3) << 2 | (by & 3) << 4)) & 1

    def sector_mask(self, sx, sy, sz):
        return self.alloc.get((sx, sy, sz), 0)

    def cell_mask(self, cx, cy, cz):  # 4x4x4 voxels, bit x + 4 z + 16 y
        m = 0
        for y in range(4):
            for z in range(4):
                for x in range(4):
                    if self.voxel(cx * 4 + x, cy * 4 + y, cz * 4 + z):
                        m |= 1 << (x + 4 * z + 16 * y)
        return m

    def group_mask(self, gx, gy, gz):
        m = 0
        for y in range(4):
            for z in range(4):
                for x in range(4):
                    if self.sector_mask(gx * 4 + x, gy * 4 + y, gz * 4 + z):
                        m |= 1 << (x + 4 * z + 16 * y)
        return m


def interaction_mask(idx, octant):  # GpuRenderer.cpp:193-210
    ox, oz, oy = idx & 3, (idx >> 2) & 3, idx >> 4
    sx, sy, sz = (1 if octant & 1 else -1), (1 if octant & 2 else -1), (1 if octant & 4 else -1)
    m = 0
    for j in range(64):
        x, y, z = ox + (j & 3) * sx, oy + (j >> 4) * sy, oz + ((j >> 2) & 3) * sz
        if 0 <= x < 4 and 0 <= y < 4 and 0 <= z < 4:
            m |= 1 << (x + 4 * z + 16 * y)
    return m


def _floor_i(v):
    f = np.floor(v)
    return int(f) if np.isfinite(f) and -2147483648.0 <= f < 2147483648.0 else -(1 << 31)


def _wrap(i):  # int32 wrap-around like the shader's ivec3 arithmetic
    return (i + (1 << 31)) % (1 << 32) - (1 << 31)


def _gmin(a, b):
    return b if b < a else a


def _gmax(a, b):
    return b if a < b else a


def cast(world: DenseWorld, o, d, wo, coarse_mode=False, aniso=False):
    """-> dict(hit, inb, capped, iters, voxel, dist, pos, normal, uv)"""
    with np.errstate(all="ignore"):
        o = [F(v) for v in o]
        d = [F(v) for v in d]
        inv = [F(1) / v for v in d]
        # clipRayToAABB(origin, dir, -wo + 1, grid - wo - 1)
        t1, t2 = [], []
        for a in range(3):
            lo = F(_wrap(1 - wo[a]))
            hi = (F(world.grid[a]) - F(wo[a])) - F(1)
            a1, a2 = (lo - o[a]) * inv[a], (hi - o[a]) * inv[a]
            t1.append(_gmin(a1, a2))
            t2.append(_gmax(a1, a2))
        tn = _gmax(t1[0], _gmax(t1[1], t1[2]))
        tf = _gmin(t2[0], _gmin(t2[1], t2[2]))
        start = [o[a] + d[a] * tn for a in range(3)] if (tn > 0 and tn < tf) else list(o)
        if coarse_mode:
            o = list(start)
        ts = [((F(0) if d[a] < 0 else F(1)) - o[a]) * inv[a] for a in range(3)]
        p = [_wrap(wo[a] + _floor_i(start[a])) for a in range(3)]
        cap = 96 if coarse_mode else 256
        octant = (0 if d[0] < 0 else 1) + (0 if d[1] < 0 else 2) + (0 if d[2] < 0 else 4)
        hit, inb, i = False, True, 0
        sd, tmin, cur = [F(0)] * 3, F(0), [F(0)] * 3
        while i < cap:
            sd = [ts[a] + F(_wrap(p[a] - wo[a])) * inv[a] for a in range(3)]
            tmin = _gmin(_gmin(sd[0], sd[1]), sd[2])
            if coarse_mode:
                tmin = tmin + F(0.001)
            elif tmin == tmin:
                tmin = (np.array([tmin], F).view(np.uint32) + np.uint32(5)).view(F)[0]
            cur = [o[a] + tmin * d[a] for a in range(3)]
            p = [_wrap(wo[a] + _floor_i(cur[a])) for a in range(3)]
            inb = 0 <= p[0] < world.grid[0] and 0 <= p[2] < world.grid[2] and 0 <= p[1] < world.grid[1]
            if not inb:
                break
            # getStepPos
            x, y, z = p
            if world.brick_allocated(x >> 3, y >> 3, z >> 3):
                mask, idx, scale = world.cell_mask(x >> 2, y >> 2, z >> 2), (x & 3) | (z & 3) << 2 | (y & 3) << 4, 1
                if (mask >> idx) & 1:
                    hit = True
                    break
            elif world.sector_mask(x >> 5, y >> 5, z >> 5) == 0:
                mask, idx, scale = world.group_mask(x >> 7, y >> 7, z >> 7), ((x >> 5) & 3) | ((z >> 5) & 3) << 2 | ((y >> 5) & 3) << 4, 32
            else:
                mask, idx, scale = world.sector_mask(x >> 5, y >> 5, z >> 5), ((x >> 3) & 3) | ((z >> 3) & 3) << 2 | ((y >> 3) & 3) << 4, 8
            if aniso:
                mask &= interaction_mask(idx, octant)
            if mask == 0:
                lod = 4
            else:  # getIsotropicLod: the 2x2x2 block of the mask that holds idx
                bx, bz, by = idx & 2, (idx >> 2) & 2, (idx >> 4) & 2
                block = sum(1 << ((bx + i2) + 4 * (bz + k2) + 16 * (by + j2)) for i2 in range(2) for j2 in range(2) for k2 in range(2))
                lod = 2 if (mask & block) == 0 else 1
            lod *= scale
            if coarse_mode and i > 30 and lod < 4:
                bit = (mask & -mask).bit_length() - 1
                p = [(x & ~3) | (bit & 3), (y & ~3) | (bit >> 4), (z & ~3) | ((bit >> 2) & 3)]
                hit = True
                break
            cm = lod - 1
            p = [(p[a] & ~cm) if d[a] < 0 else (p[a] | cm) for a in range(3)]
            i += 1
        capped = i >= cap
        side = [bool(tmin >= sd[a]) for a in range(3)]
        normal = [(0 if not side[a] else (-1 if d[a] > 0 else (1 if d[a] < 0 else 0))) for a in range(3)]
        fu, fv = (cur[1] if side[0] else cur[0]), (cur[1] if side[2] else cur[2])
        return {
            "hit": hit, "inb": inb, "capped": capped, "iters": cap if capped else i, "voxel": p, "dist": tmin, "pos": cur,
            "normal": normal if hit else [0, 0, 0], "uv": [fu - np.floor(fu), fv - np.floor(fv)] if hit else [F(0), F(0)],
            "material": int(world.palette[world.voxel(*p)]) & 0xFFFFFFFF if hit else 0,
        }
