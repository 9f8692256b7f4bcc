"""Dynamic-edit workload (INPUT generation for BASELINE.json configs[4]; not the product path, not the oracle).

The reference edits its world through VoxelMap::Set / brushes (VoxelRT/VoxelMap.h:216-264, Brush.cpp), which mark
(sector, brick) pairs in VoxelMap::DirtyLocs; the renderer's SyncBuffers then uploads exactly those bricks
(CpuRenderer.cpp:33-61, GpuRenderer.cpp:45-79).  `EditableWorld` is that host-side world for the bench and the
tests: a seeded stream of single-voxel edits (uniform positions inside a box, half of them set a voxel, half
clear one) is applied to a copy of a scene, and every frame's dirty bricks come back as sync records
`(sx, sy, sz, alloc_mask, dirty_mask, bricks[popcount(dirty & alloc), 512])` — the VrtDirtySector contract.
Like VoxelMap::Set, writing into a missing brick allocates it, and a brick whose last voxel is cleared stays
allocated; bricks are only freed by the region GC (VoxelMap.h:253-262), which the brush strokes below run like the reference does.
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
# material 0 erases.  (VoxelMap::RegionDispatchSIMD creates the bricks of the stroke's box unless it erases and deletes the ones that end
# up empty, VoxelMap.h:223-263.)
# ---------------------------------------------------------------------------------------------------------------------
_brush_lib = None


def _brushlib():
    global _brush_lib
    if _brush_lib is None:
        import ctypes as C
        from pathlib import Path

        lib = C.CDLL(str(Path(__file__).resolve().parent / "_ref" / "libbrush.so"))
        lib.brush_brick_mask.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        lib.brush_brick_mask.restype = C.c_int
        _brush_lib = lib
    return _brush_lib


def brush_dispatch(world: EditableWorld, point_a, point_b, radius=30.0, material=255, action="fill"):
    """BrushSession::Dispatch (Brush.cpp:10-37) over VoxelMap::RegionDispatchSIMD (VoxelMap.h:223-263) on an EditableWorld -> the sync
    records of everything the stroke marked dirty.  Step by step like the reference: every brick of the stroke's bounding box is visited
    (created on the way: always inside an existing sector — VoxelMap::GetBrick's quirk — and together with their sector unless the stroke
    erases; a Replace stroke's new bricks stay empty); the capsule test runs in the
    reference's arithmetic (scenes/brush.cpp); a visited brick is dirty if some voxel was inside the final mask or if it is empty afterwards;
    and visited bricks that end up empty are deleted again, with their sector if nothing else is left in it (the region's garbage
    collection).  tests/test_ref_brush_pin.py compares voxels, allocation masks and dirty bricks with the reference's own code."""
    import ctypes as C

    lib = _brushlib()
    a, b = np.asarray(point_a, np.int64), np.asarray(point_b, np.int64)
    a3, b3 = (C.c_int32 * 3)(*[int(v) for v in a]), (C.c_int32 * 3)(*[int(v) for v in b])
    pad = int(radius + 0.5)
    lo, hi = np.minimum(a, b) - pad, np.maximum(a, b) + pad  # Brush.cpp:11-12
    erasing = material == 0
    create_empty = not erasing  # :17
    replace = action == "replace"
    dirty, emptied = {}, {}
    inside = np.zeros(512, np.uint8)
    for by in range(int(lo[1]) >> 3, (int(hi[1]) >> 3) + 1):
        for bz in range(int(lo[2]) >> 3, (int(hi[2]) >> 3) + 1):
            for bx in range(int(lo[0]) >> 3, (int(hi[0]) >> 3) + 1):
                key = (bx >> 2, by >> 2, bz >> 2)
                if min(key) < 0:
                    continue  # (the reference's world is signed; nothing below zero is ever in a renderer's view)
                bi = (bx & 3) | ((bz & 3) << 2) | ((by & 3) << 4)
                d = world.sectors.get(key)
                if d is None:
                    if not create_empty:
                        continue  # GetBrick(pos, false) == nullptr for a missing SECTOR only ...
                    d = world.sectors[key] = {}
                if bi not in d:  # ... inside an existing sector the brick is created on lookup whatever `create` says (VoxelMap.cpp:122, quirk Q5)
                    d[bi] = np.zeros(512, np.uint8)
                    world._owned.add((key, bi))
                vox = d[bi]
                m = None
                if lib.brush_brick_mask(a3, b3, float(radius), bx, by, bz, inside.ctypes.data):
                    m = inside.astype(bool)
                    if replace:
                        m &= vox != 0  # :25-27
                changed = m is not None and bool(m.any())  # DispatchSIMD's result: simd::any(mask) of some invocation
                if changed:
                    if (key, bi) not in world._owned:
                        vox = d[bi] = vox.copy()
                        world._owned.add((key, bi))
                    vox[m] = material
                is_empty = not vox.any()
                if changed or is_empty:
                    dirty[key] = dirty.get(key, 0) | (1 << bi)
                    if is_empty:
                        emptied[key] = emptied.get(key, 0) | (1 << bi)
    for key, emask in emptied.items():  # VoxelMap.h:253-262
        d = world.sectors[key]
        if world.alloc_mask(key) & ~emask:
            for bi in [x for x in d if (emask >> x) & 1]:
                del d[bi]
                world._owned.discard((key, bi))
        else:
            for bi in list(d):
                world._owned.discard((key, bi))
            del world.sectors[key]
    recs = []
    for key in sorted(dirty):
        d = world.sectors.get(key)
        alloc = world.alloc_mask(key)
        dm = dirty[key]
        payload = [d[bb] for bb in sorted(d) if (dm >> bb) & 1] if d else []
        recs.append((key[0], key[1], key[2], alloc, dm, np.stack(payload) if payload else np.zeros((0, 512), np.uint8), d is None))
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
