"""Bundled-model scenes (INPUT generation for BASELINE.json configs[2]; not the product path, not the oracle).

sponza(size): the reference app's start-up model — `assets/models/Sponza/Sponza.gltf` voxelised into a cube of
`size` voxels at the origin (VoxelRT/Main.cpp:42-47 uses 2048) by the recipe of VoxelMap::VoxelizeModel
(VoxelRT/Voxelize.cpp:77-147):

  1. every base-colour texture is reduced to mip 2 (4x4 box filter) and its opaque texels (alpha >= 200) feed an
     octree colour quantiser that is cut down to <= 240 leaves (Common/PaletteBuilder.h:18-126: 6-level octree,
     least-populated parents merged first); palette entry i = mean colour of leaf i;
  2. triangles are scaled so the model's longest axis spans `size` voxels, centred in x/z, resting on y = 0;
  3. conservative surface voxelisation (Schwarz & Seidel), colour = nearest texel of mip 2 at the barycentric
     projection of the voxel corner, alpha test at 128, voxel id = nearest palette entry (Manhattan distance).

The geometry kernel is scenes/voxelizer.c (built by scenes/Makefile into scenes/_ref/libvoxelizer.so).  The glTF
is read directly (one buffer, u16 indices, f32 POSITION / TEXCOORD_0, node TRS) instead of through assimp, the
images through PIL instead of stb_image; neither is pinned against the reference (its loaders are third-party
libraries absent here), so the result is "the reference's scene by the reference's recipe", not a bit-copy.
Parity of the TRAVERSAL does not depend on it: oracle and GPU consume the same bricks.

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
def load_mip2(path: Path):
    """RGBA8 image reduced by a 4x4 box filter (two 2x2 steps with round-to-nearest, like a mip chain)."""
    from PIL import Image

    img = np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint16)
    for _ in range(2):
        h, w = img.shape[0] & ~1, img.shape[1] & ~1
        if h < 2 or w < 2:
            break
        img = img[:h, :w]
        img = (img[0::2, 0::2] + img[1::2, 0::2] + img[0::2, 1::2] + img[1::2, 1::2] + 2) >> 2
    return img.astype(np.uint8)


class OctreePalette:
    """Octree colour quantiser after Common/PaletteBuilder.h: colours are binned by their top 6 bits per channel;
    while more than `max_colors` leaves remain, the parent with the smallest population is collapsed into a leaf."""

    LEVELS = 6

    def __init__(self):
        self.count = {}  # (level, key) -> population, key = interleaved child indices down to `level`
        self.rgb = {}    # leaf (level, key) -> [sum r, sum g, sum b]

    def add_colors(self, rgb: np.ndarray):
        rgb = np.asarray(rgb, np.uint8).reshape(-1, 3)
        if rgb.size == 0:
            return
        r, g, b = (rgb[:, 0].astype(np.uint32), rgb[:, 1].astype(np.uint32), rgb[:, 2].astype(np.uint32))
        key = np.zeros(rgb.shape[0], np.uint32)
        for level in range(self.LEVELS):  # child index = r bit | g bit << 1 | b bit << 2 (PaletteBuilder.h:176-180)
            child = ((r >> (7 - level)) & 1) | (((g >> (7 - level)) & 1) << 1) | (((b >> (7 - level)) & 1) << 2)
            key = (key << np.uint32(3)) | child
        uk, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
        sr = np.bincount(inv, weights=r, minlength=uk.size)
        sg = np.bincount(inv, weights=g, minlength=uk.size)
        sb = np.bincount(inv, weights=b, minlength=uk.size)
        for k, c, a0, a1, a2 in zip(uk.tolist(), cnt.tolist(), sr.tolist(), sg.tolist(), sb.tolist()):
            leaf = (self.LEVELS, k)
            acc = self.rgb.setdefault(leaf, [0, 0, 0])
            acc[0] += int(a0)
            acc[1] += int(a1)
            acc[2] += int(a2)
            for level in range(self.LEVELS + 1):
                node = (level, k >> (3 * (self.LEVELS - level)))
                self.count[node] = self.count.get(node, 0) + c

    def build(self, max_colors=240):
        import heapq

        leaves = set(self.rgb.keys())
        children = {}
        for (level, key) in self.count:
            if level > 0:
                children.setdefault((level - 1, key >> 3), []).append((level, key))
        heap = [(self.count[p], p) for p in {(lv - 1, k >> 3) for (lv, k) in leaves}]
        heapq.heapify(heap)
        while len(leaves) > max_colors and heap:
            _, node = heapq.heappop(heap)
            if node in leaves:
                continue
            kids = [c for c in children.get(node, []) if self.count.get(c, 0) > 0]
            if not kids:
                continue
            acc = [0, 0, 0]
            stack = list(kids)
            removed = 0
            while stack:  # collapse the whole subtree into `node`
                c = stack.pop()
                if c in leaves:
                    leaves.discard(c)
                    removed += 1
                    s = self.rgb.pop(c)
                    acc[0] += s[0]
                    acc[1] += s[1]
                    acc[2] += s[2]
                else:
                    stack.extend(k for k in children.get(c, []) if self.count.get(k, 0) > 0)
            if removed == 0:
                continue
            self.rgb[node] = acc
            leaves.add(node)
            if node[0] > 0:
                parent = (node[0] - 1, node[1] >> 3)
                heapq.heappush(heap, (self.count[parent], parent))
        # depth-first child order, like Octree::FindLeafs (PaletteBuilder.h:207-221)
        order = sorted(leaves, key=lambda n: n[1] << (3 * (self.LEVELS - n[0])))
        pal = np.zeros((len(order), 3), np.uint8)
        for i, n in enumerate(order):
            c = self.count[n]
            pal[i] = [self.rgb[n][0] // c, self.rgb[n][1] // c, self.rgb[n][2] // c]
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
def _voxlib():
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
    lib.vox_triangles.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint8]
    lib.vox_triangles.restype = C.c_int64
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


def voxelize_model(gltf: Path, size: int):
    """VoxelMap::VoxelizeModel(model, startPos = 0, size = `size`^3), Voxelize.cpp:77-147."""
    tris, uvs, mats, images = load_gltf(gltf)
    # 1. textures -> mip 2, palette from their opaque texels
    tex_rgba = {}
    quant = OctreePalette()
    for p in sorted({p for p in images if p is not None}):
        img = load_mip2(p)
        tex_rgba[p] = img
        if img.shape[0] * 4 <= 4 or img.shape[1] * 4 <= 4:  # Voxelize.cpp:81 skips tiny (placeholder) textures
            continue
        opaque = img[..., 3] >= 200  # :88
        quant.add_colors(img[..., :3][opaque])
    pal_rgb = quant.build(240)  # :94
    palette = terrain.reference_palette()  # debug / emissive entries of Main.cpp:52-60 stay in place
    for i, c in enumerate(pal_rgb):
        palette[i] = terrain.encode_material(int(c[0]), int(c[1]), int(c[2]))  # :96-99 (entry 0 is the empty voxel id)
    # 2. placement (:101-115): bounds over the (linear-part) transformed vertices
    flat = tris.reshape(-1, 3)
    bmin, bmax = flat.min(axis=0), flat.max(axis=0)
    rng = bmax - bmin
    scale = np.float32(size) / np.float32(rng.max())
    center = (np.float32(size) - rng * scale) * np.float32(0.5)
    center[1] = 0
    vt = ((tris - bmin) * scale + center).astype(np.float32)
    # 3. quantise every texture once (texel -> palette index, 255 = transparent), then rasterise
    tex_list = sorted(tex_rgba.keys())
    tex_index = {p: i for i, p in enumerate(tex_list)}
    quantised = []
    for p in tex_list:
        img = tex_rgba[p]
        idx = nearest_palette_index(pal_rgb, img[..., :3]).reshape(img.shape[:2])
        idx = np.where(img[..., 3] >= 128, idx, 255).astype(np.uint8)  # :137 alpha test
        quantised.append(np.ascontiguousarray(idx))
    tri_tex = np.array([tex_index[images[m]] if (0 <= m < len(images) and images[m] is not None) else -1 for m in mats.tolist()], np.int32)
    lib = _voxlib()
    nb = size // 8
    grid = lib.vox_create(nb, nb, nb)
    if not grid:
        raise MemoryError("voxeliser grid")
    try:
        ptrs = (C.c_void_p * max(1, len(quantised)))(*[q.ctypes.data for q in quantised])
        tw = np.array([q.shape[1] for q in quantised] or [1], np.int32)
        th = np.array([q.shape[0] for q in quantised] or [1], np.int32)
        vt = np.ascontiguousarray(vt)
        uvc = np.ascontiguousarray(uvs, dtype=np.float32)
        degenerate = lib.vox_triangles(grid, vt.shape[0], vt.ctypes.data, uvc.ctypes.data, tri_tex.ctypes.data, ptrs, tw.ctypes.data, th.ctypes.data, 1)
        n = lib.vox_brick_count(grid)
        bricks = np.zeros((n, 512), np.uint8)
        coords = np.zeros((n, 3), np.int32)
        lib.vox_export(grid, bricks.ctypes.data, coords.ctypes.data)
        written = lib.vox_voxels_set(grid)
    finally:
        lib.vox_destroy(grid)
    scene = bricks_to_scene(coords, bricks, palette, f"{gltf.parent.name}/{gltf.name} voxelised into {size}^3 ({vt.shape[0]} triangles, {degenerate} degenerate, {written} voxel writes)")
    return scene


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
