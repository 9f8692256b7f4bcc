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
