"""Bundled-model scenes (INPUT generation for BASELINE.json configs[2]; not the product path, not the oracle).

sponza(size): the reference app's start-up model — `assets/models/Sponza/Sponza.gltf` voxelised into a cube of
`size` voxels at the origin (VoxelRT/Main.cpp:42-47 uses 2048) by the recipe of VoxelMap::VoxelizeModel
(VoxelRT/Voxelize.cpp:77-147):

  1. every base-colour texture gets the sampler's mip chain (2x2 fp32 averages) and the opaque texels (alpha >= 200)
     of level 2 feed an octree colour quantiser that is cut down to <= 240 leaves (Common/PaletteBuilder.h:18-62:
     6-level octree, least-populated parents folded first); palette entry i = mean colour of leaf i;
  2. triangles are scaled so the model's longest axis spans `size` voxels, centred in x/z, resting on y = 0;
  3. conservative surface voxelisation (Schwarz & Seidel), colour = the sampler's bilinear LEVEL-0 tap at the
     barycentric projection of the voxel corner (the `2` handed to Sample() there is overridden by the UV
     derivatives, which are zero for a broadcast coordinate, Texture.h:504-509), alpha test at 128, voxel id =
     nearest palette entry (Manhattan distance).

Steps 1-3 are PINNED: tests/test_ref_pin.py hands the same decoded arrays to the reference's own VoxelizeModel
(oracle/_ref, compiled from Voxelize.cpp + PaletteBuilder.h where they lie) and demands the same palette and the same
voxels.  What stays unpinned is DECODING: the glTF is read directly (one buffer, u16 indices, f32 POSITION /
TEXCOORD_0, node TRS) instead of through assimp, the images through PIL instead of stb_image — third-party loaders
absent here.  The geometry kernel is scenes/voxelizer.c (built by scenes/Makefile into scenes/_ref/libvoxelizer.so).
Parity of the TRAVERSAL does not depend on any of it: oracle and GPU consume the same bricks.

The voxelised scene is cached as scenes/_ref/sponza_<size>.dat in the reference's own "cvox 0004" cache format
(scenes/cvox.py — the file the reference app itself would load as logs/voxels_2k_sponza.dat, Main.cpp:38-49); git-ignored,
travels to the GPU box, where /root/reference does not exist.
"""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path

import numpy as np

from . import terrain

HERE = Path(__file__).resolve().parent
REF_MODEL = Path("/root/reference/assets/models/Sponza/Sponza.gltf")
VOX_LIB = HERE / "_ref" / "libvoxelizer.so"

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_WIDTH = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


# ---------------------------------------------------------------------------------------------
# glTF 2.0 (the subset the bundled models use)
# ---------------------------------------------------------------------------------------------
def _accessor(g, buffers, idx):
    a = g["accessors"][idx]
    bv = g["bufferViews"][a["bufferView"]]
    dt = np.dtype(_COMPONENT[a["componentType"]])
    w = _WIDTH[a["type"]]
    start = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
    stride = bv.get("byteStride", 0) or dt.itemsize * w
    buf = buffers[bv["buffer"]]
    elem = dt.itemsize * w
    if a["count"] == 0:
        return np.zeros((0, w), dt)
    raw = np.frombuffer(buf, dtype=np.uint8, offset=start, count=stride * (a["count"] - 1) + elem)
    rows = np.lib.stride_tricks.as_strided(raw, shape=(a["count"], elem), strides=(stride, 1))
    return np.ascontiguousarray(rows).view(dt).reshape(a["count"], w)


def _node_matrix(n):
    if "matrix" in n:
        return np.array(n["matrix"], np.float64).reshape(4, 4).T
    m = np.eye(4)
    if "scale" in n:
        m = np.diag(list(n["scale"]) + [1.0]) @ m
    if "rotation" in n:
        x, y, z, w = n["rotation"]
        r = np.array(
            [
                [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 0],
                [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 0],
                [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y), 0],
                [0, 0, 0, 1],
            ]
        )
        m = r @ m
    if "translation" in n:
        t = np.eye(4)
        t[:3, 3] = n["translation"]
        m = t @ m
    return m


def load_gltf(path: Path):
    """-> (tris float32[T,3,3] in model space, uvs float32[T,3,2], material int32[T], image path per material or None).

    Like VoxelizeModel, positions are transformed by the LINEAR part of the node matrices only (w = 0,
    Voxelize.cpp:103,126)."""
    path = Path(path)
    g = json.loads(path.read_text())
    buffers = [(path.parent / b["uri"]).read_bytes() for b in g["buffers"]]
    tris, uvs, mats = [], [], []

    def visit(ni, parent):
        n = g["nodes"][ni]
        m = parent @ _node_matrix(n)
        if "mesh" in n:
            for prim in g["meshes"][n["mesh"]]["primitives"]:
                if prim.get("mode", 4) != 4:
                    continue
                pos = _accessor(g, buffers, prim["attributes"]["POSITION"]).astype(np.float32)
                if "TEXCOORD_0" in prim["attributes"]:
                    uv = _accessor(g, buffers, prim["attributes"]["TEXCOORD_0"]).astype(np.float32)
                else:
                    uv = np.zeros((pos.shape[0], 2), np.float32)
                idx = _accessor(g, buffers, prim["indices"]).reshape(-1).astype(np.int64) if "indices" in prim else np.arange(pos.shape[0])
                idx = idx[: (idx.size // 3) * 3].reshape(-1, 3)
                p = (pos.astype(np.float64) @ m[:3, :3].T).astype(np.float32)
                tris.append(p[idx])
                uvs.append(uv[idx])
                mats.append(np.full(idx.shape[0], prim.get("material", -1), np.int32))
        for c in n.get("children", []):
            visit(c, m)

    scene = g["scenes"][g.get("scene", 0)]
    for ni in scene["nodes"]:
        visit(ni, np.eye(4))
    images = []
    for mat in g.get("materials", []):
        tex = mat.get("pbrMetallicRoughness", {}).get("baseColorTexture")
        if tex is None:
            images.append(None)
            continue
        src = g["textures"][tex["index"]]["source"]
        images.append(path.parent / g["images"][src]["uri"])
    return np.concatenate(tris), np.concatenate(uvs), np.concatenate(mats), images


# ---------------------------------------------------------------------------------------------
# textures and palette
# ---------------------------------------------------------------------------------------------
def load_rgba(path: Path):
    """Decoded image as packed RGBA8 words (R in the low byte), uint32[h, w] — what stbi_load(..., 4) hands Scene.cpp:102."""
    from PIL import Image

    img = np.ascontiguousarray(np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8))
    return img.view(np.uint32).reshape(img.shape[0], img.shape[1])


def mip_chain(level0: np.ndarray, max_levels=8):
    """Texture2D's mip chain (Texture.h:403-419, 683-704): level i exists while both sides >> i are >= 4; 2x2 fp32 average."""
    h, w = level0.shape
    if w & (w - 1) or h & (h - 1):
        raise ValueError(f"texture {w}x{h}: swr::Texture2D takes power-of-two sizes only (Texture.h:404)")
    lib = _voxlib()
    levels = [np.ascontiguousarray(level0, dtype=np.uint32)]
    while len(levels) < max_levels and (w >> len(levels)) >= 4 and (h >> len(levels)) >= 4:
        src = levels[-1]
        dst = np.empty((src.shape[0] // 2, src.shape[1] // 2), np.uint32)
        lib.vox_mip_half(src.ctypes.data, src.shape[1], src.shape[0], dst.ctypes.data)
        levels.append(dst)
    return levels


def palette_texels(levels):
    """The texels VoxelizeModel's palette pass reads from one texture (Voxelize.cpp:80-92): IterateTiles over (w/4, h/4) in 4x4
    lane tiles, u = (x + 0.5) / (w/4), nearest texel of the level the 4-texel UV step selects (2, or the last level of a texture
    too small to have one), Repeat addressing — lanes past w/4 in a tile wrap around and count again, like they do there."""
    h, w = levels[0].shape
    level = min(2, len(levels) - 1)

    def taps(n):
        cells = n // 4
        lanes = np.arange(-(-cells // 4) * 4, dtype=np.int32)
        u = (lanes.astype(np.float32) + np.float32(0.5)) * (np.float32(1.0) / np.float32(cells))
        ix = np.rint(u * np.float32(n << 8)).astype(np.int64).astype(np.int32) & ((n << 8) - 1)
        return (ix >> level) >> 8

    tx, ty = taps(w), taps(h)
    return levels[level][ty[:, None], tx[None, :]].reshape(-1)


class OctreePalette:
    """Octree colour quantiser, PaletteBuilder.h:12-62,150-240 step by step: colours are binned by their top 6 bits per channel; the
    candidate set holds PARENTS of leaves ordered by (population, larger storage index first); while the running leaf count exceeds
    `max_colors` the first candidate is folded into a leaf.  The running count drops by (direct children - 1) per fold even when a
    child was itself an unfolded subtree (Reduce's recursive return value is dropped, :225-231), so the cut goes deeper than 240
    whenever that happens — kept, the palette is the reference's."""

    LEVELS = 6

    def __init__(self):
        self.count = {}  # storage index -> population (root = 0, children of i at ((i + 1) << 3) + c, PaletteBuilder.h:182-193)
        self.rgb = {}    # storage index of a leaf -> [sum r, sum g, sum b]

    def add_colors(self, rgb: np.ndarray):
        rgb = np.asarray(rgb, np.uint8).reshape(-1, 3)
        if rgb.size == 0:
            return
        r, g, b = (rgb[:, 0].astype(np.int64), rgb[:, 1].astype(np.int64), rgb[:, 2].astype(np.int64))
        key = np.zeros(rgb.shape[0], np.int64)
        for level in range(self.LEVELS):  # child index = r bit | g bit << 1 | b bit << 2 (PaletteBuilder.h:176-180)
            child = ((r >> (7 - level)) & 1) | (((g >> (7 - level)) & 1) << 1) | (((b >> (7 - level)) & 1) << 2)
            key = (key << 3) | child
        uk, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        sr = np.bincount(inv, weights=r, minlength=uk.size)
        sg = np.bincount(inv, weights=g, minlength=uk.size)
        sb = np.bincount(inv, weights=b, minlength=uk.size)
        for k, c, a0, a1, a2 in zip(uk.tolist(), cnt.tolist(), sr.tolist(), sg.tolist(), sb.tolist()):
            node = 0
            self.count[0] = self.count.get(0, 0) + c
            for level in range(self.LEVELS):
                node = ((node + 1) << 3) + ((k >> (3 * (self.LEVELS - 1 - level))) & 7)
                self.count[node] = self.count.get(node, 0) + c
            acc = self.rgb.setdefault(node, [0, 0, 0])
            acc[0] += int(a0)
            acc[1] += int(a1)
            acc[2] += int(a2)

    def build(self, max_colors=240):
        import heapq

        count = self.count
        leaf = set(self.rgb.keys())
        rgb = {k: list(v) for k, v in self.rgb.items()}

        def parent(i):
            return (i >> 3) - 1

        def kids(i):
            base = (i + 1) << 3
            return [base + c for c in range(8) if count.get(base + c, 0) > 0]

        def reduce(i):  # Octree::Reduce
            acc = rgb.setdefault(i, [0, 0, 0])
            n = 0
            for c in kids(i):
                if c not in leaf:
                    reduce(c)
                for j in range(3):
                    acc[j] += rgb[c][j]
                n += 1
            leaf.add(i)
            return n - 1

        def find_leafs(i=0, out=None):  # Octree::FindLeafs: depth first, children in index order
            out = [] if out is None else out
            if i in leaf:
                out.append(i)
                return out
            for c in kids(i):
                find_leafs(c, out)
            return out

        num = len(leaf)
        members = {parent(i) for i in leaf}
        heap = [(count[i], -i) for i in members]  # std::set ordered by Count, ties: larger address first (:28)
        heapq.heapify(heap)
        while num > max_colors and heap:
            _, neg = heapq.heappop(heap)
            node = -neg
            members.discard(node)
            if node in leaf:
                continue
            num -= reduce(node)
            if node == 0:
                break  # the reference would index the root's parent here; 240 leaves are never reached from a single node
            par = parent(node)
            if par not in members:
                members.add(par)
                heapq.heappush(heap, (count[par], -par))
        order = find_leafs()
        pal = np.zeros((len(order), 3), np.uint8)
        for i, n in enumerate(order):
            c = count[n]
            pal[i] = [rgb[n][0] // c, rgb[n][1] // c, rgb[n][2] // c]
        return pal


def nearest_palette_index(pal: np.ndarray, rgb: np.ndarray, chunk=1 << 16):
    """PaletteBuilder::FindIndex (PaletteBuilder.h:66-128): smallest Manhattan distance, first entry wins ties."""
    rgb = np.asarray(rgb, np.int16).reshape(-1, 3)
    out = np.empty(rgb.shape[0], np.uint8)
    p = pal.astype(np.int16)
    for s in range(0, rgb.shape[0], chunk):
        d = np.abs(rgb[s : s + chunk, None, :] - p[None, :, :]).sum(axis=2)
        out[s : s + chunk] = np.argmin(d, axis=1).astype(np.uint8)
    return out


# ---------------------------------------------------------------------------------------------
# voxelisation
# ---------------------------------------------------------------------------------------------
_vox = None


def _voxlib():
    global _vox
    if _vox is not None:
        return _vox
    lib = C.CDLL(str(VOX_LIB))
    lib.vox_create.argtypes = [C.c_int32] * 3
    lib.vox_create.restype = C.c_void_p
    lib.vox_destroy.argtypes = [C.c_void_p]
    lib.vox_destroy.restype = None
    lib.vox_brick_count.argtypes = [C.c_void_p]
    lib.vox_brick_count.restype = C.c_int64
    lib.vox_voxels_set.argtypes = [C.c_void_p]
    lib.vox_voxels_set.restype = C.c_int64
    lib.vox_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vox_export.restype = None
    lib.vox_triangles.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_uint8]
    lib.vox_triangles.restype = C.c_int64
    lib.vox_mip_half.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.vox_mip_half.restype = None
    _vox = lib
    return lib


def bricks_to_scene(coords: np.ndarray, bricks: np.ndarray, palette: np.ndarray, name: str):
    """(brick coordinates [n,3], bricks [n,512]) -> scene dict (non-empty bricks only, ascending brick index per sector)."""
    keep = bricks.any(axis=1)
    coords, bricks = coords[keep], bricks[keep]
    sec = coords >> 2
    bi = (coords[:, 0] & 3) | ((coords[:, 2] & 3) << 2) | ((coords[:, 1] & 3) << 4)
    order = np.lexsort((bi, sec[:, 0], sec[:, 2], sec[:, 1]))
    sec, bi, bricks = sec[order], bi[order], bricks[order]
    sectors = {}
    if sec.shape[0]:
        change = np.nonzero(np.any(sec[1:] != sec[:-1], axis=1))[0] + 1
        starts = np.concatenate([[0], change])
        ends = np.concatenate([change, [sec.shape[0]]])
        for s, e in zip(starts.tolist(), ends.tolist()):
            mask = 0
            for b in bi[s:e].tolist():
                mask |= 1 << b
            sectors[(int(sec[s, 0]), int(sec[s, 1]), int(sec[s, 2]))] = (mask, bricks[s:e])
    return {"sectors": sectors, "palette": palette, "name": name}


FLT_MIN = np.float32(1.17549435e-38)


def voxelize_arrays(tris, uvs, tri_tex, textures, size: int, name="model", grid_bricks=None):
    """VoxelMap::VoxelizeModel(model, startPos = 0, size = `size`^3), Voxelize.cpp:77-147, on a decoded model: triangles float32[T,3,3]
    (node transforms applied), uvs float32[T,3,2], a texture id per triangle (-1: none) and the level-0 RGBA8 images (uint32[h,w]).
    -> (scene dict, palette colours uint8[n,3]).  tests/test_ref_pin.py runs the reference's own VoxelizeModel on the same arrays and
    demands the same palette and the same voxels."""
    tris = np.ascontiguousarray(tris, np.float32)
    # 1. palette from the opaque level-2 texels of every texture larger than 4x4 (:80-94)
    chains = [mip_chain(t) for t in textures]
    quant = OctreePalette()
    for levels in chains:
        h, w = levels[0].shape
        if w <= 4 or h <= 4:  # :81 (placeholder textures)
            continue
        tex = palette_texels(levels)
        tex = tex[(tex >> 24) >= 200]  # :88
        quant.add_colors(np.stack([tex & 255, (tex >> 8) & 255, (tex >> 16) & 255], axis=1).astype(np.uint8))
    pal_rgb = quant.build(240)  # :94
    palette = terrain.reference_palette()  # debug / emissive entries of Main.cpp:52-60 stay in place
    for i, c in enumerate(pal_rgb):
        palette[i] = terrain.encode_material(int(c[0]), int(c[1]), int(c[2]))  # :96-99 (entry 0 is the empty voxel id)
    # 2. placement (:101-115).  The running minimum starts at +FLT_MIN (sic, :101), so a model that lies wholly on the positive side
    #    of an axis is measured from ~0 there, not from its own minimum.
    flat = tris.reshape(-1, 3)
    bmin = np.minimum(FLT_MIN, flat.min(axis=0)).astype(np.float32)
    bmax = np.maximum(np.float32(-3.4028235e38), flat.max(axis=0)).astype(np.float32)
    rng = bmax - bmin
    scale = np.float32(size) / np.float32(rng.max())
    center = (np.float32(size) - rng * scale) * np.float32(0.5)
    center[1] = 0
    vt = np.ascontiguousarray(((tris - bmin) * scale + center).astype(np.float32))
    # 3. rasterise; colour = the sampler's level-0 bilinear tap, alpha test, nearest palette entry (:117-146)
    lib = _voxlib()
    nb = size // 8 if grid_bricks is None else grid_bricks
    grid = lib.vox_create(nb, nb, nb)
    if not grid:
        raise MemoryError("voxeliser grid")
    try:
        imgs = [c[0] for c in chains]
        ptrs = (C.c_void_p * max(1, len(imgs)))(*[q.ctypes.data for q in imgs])
        tw = np.array([q.shape[1] for q in imgs] or [1], np.int32)
        th = np.array([q.shape[0] for q in imgs] or [1], np.int32)
        uvc = np.ascontiguousarray(uvs, dtype=np.float32)
        tri_tex = np.ascontiguousarray(tri_tex, np.int32)
        palc = np.ascontiguousarray(pal_rgb, np.uint8)
        degenerate = lib.vox_triangles(grid, vt.shape[0], vt.ctypes.data, uvc.ctypes.data, tri_tex.ctypes.data, ptrs, tw.ctypes.data, th.ctypes.data,
                                       palc.ctypes.data, palc.shape[0], 1)
        if degenerate < 0:
            raise MemoryError("voxeliser colour memo")
        n = lib.vox_brick_count(grid)
        bricks = np.zeros((n, 512), np.uint8)
        coords = np.zeros((n, 3), np.int32)
        lib.vox_export(grid, bricks.ctypes.data, coords.ctypes.data)
        written = lib.vox_voxels_set(grid)
    finally:
        lib.vox_destroy(grid)
    scene = bricks_to_scene(coords, bricks, palette, f"{name} voxelised into {size}^3 ({vt.shape[0]} triangles, {degenerate} zero-area, {written} voxel writes)")
    return scene, pal_rgb


def voxelize_model(gltf: Path, size: int):
    """The bundled glTF through voxelize_arrays."""
    tris, uvs, mats, images = load_gltf(gltf)
    tex_list = sorted({p for p in images if p is not None})
    tex_index = {p: i for i, p in enumerate(tex_list)}
    textures = [load_rgba(p) for p in tex_list]
    tri_tex = np.array([tex_index[images[m]] if (0 <= m < len(images) and images[m] is not None) else -1 for m in mats.tolist()], np.int32)
    return voxelize_arrays(tris, uvs, tri_tex, textures, size, name=f"{gltf.parent.name}/{gltf.name}")[0]


def _cache_path(size):
    return HERE / "_ref" / f"sponza_{size}.dat"


def sponza_available(size=2048):
    return _cache_path(size).exists() or (REF_MODEL.exists() and VOX_LIB.exists())


def sponza(size=2048):
    """The reference app's model scene (Main.cpp:42-47).  Cached; the cache file is what the GPU box sees."""
    from . import cvox

    cp = _cache_path(size)
    if cp.exists():
        scene = cvox.load_cvox(cp)
        scene["name"] = f"Sponza/Sponza.gltf voxelised into {size}^3 ({cp.name})"
        return scene
    if not (REF_MODEL.exists() and VOX_LIB.exists()):
        raise FileNotFoundError("Sponza scene: neither the cache (scenes/_ref) nor the reference assets + voxeliser are present")
    scene = voxelize_model(REF_MODEL, size)
    cp.parent.mkdir(parents=True, exist_ok=True)
    cvox.save_cvox(scene, cp)
    return scene


if __name__ == "__main__":
    import sys
    import time

    for size in [int(a) for a in sys.argv[1:]] or [512]:
        t0 = time.time()
        sc = sponza(size)
        st = terrain.scene_stats(sc)
        print(f"sponza {size}: {st}  ({time.time() - t0:.1f} s)  {sc['name']}")
