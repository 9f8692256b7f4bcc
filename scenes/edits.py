"""Dynamic-edit workload (INPUT generation for BASELINE.json configs[4]; not the product path, not the oracle).

The reference edits its world through VoxelMap::Set / brushes (VoxelRT/VoxelMap.h:216-264, Brush.cpp), which mark
(sector, brick) pairs in VoxelMap::DirtyLocs; the renderer's SyncBuffers then uploads exactly those bricks
(CpuRenderer.cpp:33-61, GpuRenderer.cpp:45-79).  `EditableWorld` is that host-side world for the bench and the
tests: a seeded stream of single-voxel edits (uniform positions inside a box, half of them set a voxel, half
clear one) is applied to a copy of a scene, and every frame's dirty bricks come back as sync records
`(sx, sy, sz, alloc_mask, dirty_mask, bricks[popcount(dirty & alloc), 512])` — the VrtDirtySector contract.
Like VoxelMap::Set, writing into a missing brick allocates it, and a brick whose last voxel is cleared stays
allocated (bricks are only freed by the region GC, VoxelMap.h:254-262).
"""
from __future__ import annotations

import numpy as np


class EditableWorld:
    def __init__(self, scene):
        self.sectors = {}  # key -> {brick index: uint8[512]} (bricks are copied on first write)
        self._owned = set()
        for key, (mask, bricks) in scene["sectors"].items():
            d = {}
            i = 0
            for b in range(64):
                if (mask >> b) & 1:
                    d[b] = bricks[i]
                    i += 1
            self.sectors[key] = d

    def alloc_mask(self, key):
        m = 0
        for b in self.sectors.get(key, {}):
            m |= 1 << b
        return m

    def apply(self, pos: np.ndarray, ids: np.ndarray):
        """pos int[n,3] voxel coordinates, ids uint8[n] (0 = clear).  Returns the frame's sync records."""
        pos = np.asarray(pos, np.int64)
        sec = pos >> 5
        bidx = ((pos[:, 0] >> 3) & 3) | (((pos[:, 2] >> 3) & 3) << 2) | (((pos[:, 1] >> 3) & 3) << 4)
        vidx = (pos[:, 0] & 7) | ((pos[:, 2] & 7) << 3) | ((pos[:, 1] & 7) << 6)
        dirty = {}
        for i in range(pos.shape[0]):
            key = (int(sec[i, 0]), int(sec[i, 1]), int(sec[i, 2]))
            b = int(bidx[i])
            v = int(ids[i])
            d = self.sectors.get(key)
            if d is None or b not in d:
                if v == 0:
                    continue  # clearing air: Set() on a missing brick with an empty voxel changes nothing
                if d is None:
                    d = self.sectors[key] = {}
                d[b] = np.zeros(512, np.uint8)
                self._owned.add((key, b))
            elif (key, b) not in self._owned:
                d[b] = d[b].copy()
                self._owned.add((key, b))
            d[b][int(vidx[i])] = v
            dirty[key] = dirty.get(key, 0) | (1 << b)
        recs = []
        for key in sorted(dirty):
            d = self.sectors[key]
            dm = dirty[key]
            payload = np.stack([d[b] for b in sorted(d) if (dm >> b) & 1])
            recs.append((key[0], key[1], key[2], self.alloc_mask(key), dm, payload))
        return recs

    def to_scene(self, palette, name="edited"):
        sectors = {}
        for key, d in self.sectors.items():
            if d:
                sectors[key] = (self.alloc_mask(key), np.stack([d[b] for b in sorted(d)]))
        return {"sectors": sectors, "palette": palette, "name": name}


def random_edit_frames(scene, n_frames, edits_per_frame, seed=1, box=((0, 768), (96, 224), (0, 768)), ids=(245, 246, 247, 248, 252, 255)):
    """-> (list of per-frame sync record lists, the EditableWorld after the last frame)."""
    rng = np.random.default_rng(seed)
    world = EditableWorld(scene)
    frames = []
    ids = np.asarray(ids, np.uint8)
    for _ in range(n_frames):
        pos = np.stack([rng.integers(lo, hi, edits_per_frame) for lo, hi in box], axis=1)
        val = np.where(rng.random(edits_per_frame) < 0.5, ids[rng.integers(0, ids.size, edits_per_frame)], 0).astype(np.uint8)
        frames.append(world.apply(pos, val))
    return frames, world


# ---------------------------------------------------------------------------------------------------------------------
# the reference's brush (VoxelRT/Brush.cpp:3-37, Brush.h:9-17): a capsule of radius 30 from the previous to the current
# brush position; Fill writes the material into every voxel whose CENTRE is inside, Replace only into non-empty voxels,
# material 0 erases.  (VoxelMap::RegionDispatchSIMD creates bricks only when filling, VoxelMap.h:216-264.)
# ---------------------------------------------------------------------------------------------------------------------
def _capsule_inside(px, py, pz, a, b, r):
    """sdCapsule(p, a, b, r) < 0 (Brush.cpp:4-8) for voxel centres p; float32 like the reference."""
    f = np.float32
    pa = [px - f(a[0]), py - f(a[1]), pz - f(a[2])]
    ba = [f(b[0] - a[0]), f(b[1] - a[1]), f(b[2] - a[2])]
    bb = f(ba[0] * ba[0] + ba[1] * ba[1] + ba[2] * ba[2])
    if bb > 0:
        h = np.clip((pa[0] * ba[0] + pa[1] * ba[1] + pa[2] * ba[2]) / bb, f(0), f(1))
    else:
        h = np.zeros_like(px)
    d = [pa[i] - ba[i] * h for i in range(3)]
    return np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) - f(r) < 0


def brush_dispatch(world: EditableWorld, point_a, point_b, radius=30.0, material=255, action="fill"):
    """BrushSession::Dispatch on an EditableWorld -> sync records of the bricks it changed."""
    a, b = np.asarray(point_a, np.int64), np.asarray(point_b, np.int64)
    pad = int(radius + 0.5)
    lo, hi = np.minimum(a, b) - pad, np.maximum(a, b) + pad  # Brush.cpp:11-12
    erasing = material == 0
    dirty = {}
    ax = np.arange(8, dtype=np.float32) + np.float32(0.5)
    for by in range(int(lo[1]) >> 3, (int(hi[1]) >> 3) + 1):
        for bz in range(int(lo[2]) >> 3, (int(hi[2]) >> 3) + 1):
            for bx in range(int(lo[0]) >> 3, (int(hi[0]) >> 3) + 1):
                # voxel index x | z << 3 | y << 6  ->  arrays shaped [y, z, x]
                py, pz, px = np.meshgrid(ax + np.float32(by * 8), ax + np.float32(bz * 8), ax + np.float32(bx * 8), indexing="ij")
                inside = _capsule_inside(px, py, pz, a, b, radius)
                ix, iy, iz = px.astype(np.int64), py.astype(np.int64), pz.astype(np.int64)
                inside &= (ix >= lo[0]) & (ix <= hi[0]) & (iy >= lo[1]) & (iy <= hi[1]) & (iz >= lo[2]) & (iz <= hi[2])
                if not inside.any():
                    continue
                key = (bx >> 2, by >> 2, bz >> 2)
                if min(key) < 0:
                    continue
                bi = (bx & 3) | ((bz & 3) << 2) | ((by & 3) << 4)
                d = world.sectors.get(key)
                have = d is not None and bi in d
                if not have and (erasing or action == "replace"):
                    continue  # nothing to erase / replace in a missing brick (createEmpty = false)
                if not have:
                    if d is None:
                        d = world.sectors[key] = {}
                    d[bi] = np.zeros(512, np.uint8)
                    world._owned.add((key, bi))
                elif (key, bi) not in world._owned:
                    d[bi] = d[bi].copy()
                    world._owned.add((key, bi))
                vox = d[bi].reshape(8, 8, 8)
                m = inside & (vox != 0) if (action == "replace" or erasing) else inside
                if not m.any():
                    continue
                vox[m] = material
                dirty[key] = dirty.get(key, 0) | (1 << bi)
    recs = []
    for key in sorted(dirty):
        d = world.sectors[key]
        dm = dirty[key]
        recs.append((key[0], key[1], key[2], world.alloc_mask(key), dm, np.stack([d[bb] for bb in sorted(d) if (dm >> bb) & 1])))
    return recs


def brush_stroke_frames(scene, n_frames, seed=1, radius=30.0, box=((120, 640), (100, 170), (120, 640)), step=24):
    """A seeded brush session: the brush position random-walks through `box`, `step` voxels per frame; strokes alternate between
    filling with an emissive material, erasing, and replacing (eight frames each).  -> (per-frame record lists, world)."""
    rng = np.random.default_rng(seed)
    world = EditableWorld(scene)
    pos = np.array([rng.integers(lo, hi) for lo, hi in box], np.int64)
    frames = []
    for f in range(n_frames):
        delta = rng.normal(size=3)
        delta = (delta / np.linalg.norm(delta) * step).astype(np.int64)
        nxt = np.array([int(np.clip(pos[a] + delta[a], box[a][0], box[a][1])) for a in range(3)], np.int64)
        phase = (f // 8) % 3
        material, action = ((254, "fill"), (0, "replace"), (252, "replace"))[phase]
        frames.append(brush_dispatch(world, pos, nxt, radius, material, action))
        pos = nxt
    return frames, world
