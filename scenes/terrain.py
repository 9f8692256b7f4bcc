"""Scene generators (INPUT data for tests and the bench; not the product path, not the oracle).

terrain_fastnoise(): the reference's procedural terrain, VoxelRT/TerrainGenerator.cpp:5-34, with
the density taken from the reference's own vendored FastNoise2 (scenes/_ref/libFastNoise.so,
built by scenes/Makefile from /root/reference/deps/FastNoise2) — same node tree, frequency
0.004, seed 12345, y offset -96, grass ids 245..248.  Sector range of the reference app:
x,z in [0,24), y in [0,7) (VoxelRT/Main.cpp:63-69).

terrain_hash(): an integer-only stand-in (value noise from a 32-bit hash) used when the
FastNoise2 library is unavailable and for bit-reproducible golden fixtures.

A scene is {"sectors": {(sx,sy,sz): (alloc_mask:int, bricks: uint8[k,512])}, "palette": uint64[256]}
with bricks in ascending brick-index order (brick index = bx | bz<<2 | by<<4; voxel index =
x | z<<3 | y<<6, VoxelRT/VoxelMap.h:100-102).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
FN_LIB = HERE / "_ref" / "libFastNoise.so"
CACHE_DIR = HERE / "_cache"

NODE_TREE = (
    b"EQACAAAAAAAgQBAAAAAAQBkAEwDD9Sg/DQAEAAAAAAAgQAkAAGZmJj8AAAAAPwEEAAAAAAAAAEBAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAM3MTD4AMzMzPwAAAAA/"
)  # TerrainGenerator.cpp:6


def encode_material(r, g, b, fuzz=255, emission=0.0):
    """Material::GetEncoded, VoxelRT/VoxelMap.h:27-41 (RGB565 | f16 emission << 16 | fuzz << 32)."""
    h = int(np.float16(np.float32(emission)).view(np.uint16))
    return ((r >> 3) << 11) | ((g >> 2) << 5) | (b >> 3) | (h << 16) | (fuzz << 32)


def reference_palette():
    """Debug palette of the reference app, VoxelRT/Main.cpp:52-60 (greens 245-248, emissive 252-255)."""
    pal = np.zeros(256, np.uint64)
    rng = np.random.default_rng(7)
    for i in range(1, 245):  # arbitrary matte colours for ids the terrain never uses
        c = rng.integers(40, 230, 3)
        pal[i] = encode_material(int(c[0]), int(c[1]), int(c[2]))
    pal[245] = encode_material(70, 150, 64)  # Main.cpp:52-55
    pal[246] = encode_material(110, 150, 64)
    pal[247] = encode_material(138, 160, 72)
    pal[248] = encode_material(60, 130, 56)
    pal[252] = encode_material(255, 48, 48, emission=0.8)  # Main.cpp:57-60
    pal[253] = encode_material(48, 255, 48, emission=0.8)
    pal[254] = encode_material(48, 48, 255, emission=0.8)
    pal[255] = encode_material(255, 255, 255, emission=10.0)
    return pal


def sector_to_bricks(vox_yzx: np.ndarray):
    """32^3 voxel ids indexed [y,z,x] -> (alloc_mask of NON-EMPTY bricks, bricks[k,512]).
    Only non-empty bricks are kept, as TerrainGenerator::WorkerFn does (TerrainGenerator.cpp:139-143)."""
    b = vox_yzx.reshape(4, 8, 4, 8, 4, 8).transpose(0, 2, 4, 1, 3, 5).reshape(64, 512)
    nonempty = b.any(axis=1)
    mask = 0
    for i in np.nonzero(nonempty)[0]:
        mask |= 1 << int(i)
    return mask, np.ascontiguousarray(b[nonempty])


def fastnoise_available():
    return FN_LIB.exists()


_fn = None


def _fastnoise():
    global _fn
    if _fn is None:
        lib = C.CDLL(str(FN_LIB))
        lib.fnNewFromEncodedNodeTree.argtypes = [C.c_char_p, C.c_uint]
        lib.fnNewFromEncodedNodeTree.restype = C.c_void_p
        lib.fnGenUniformGrid3D.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_float, C.c_int, C.c_void_p]
        lib.fnGenUniformGrid3D.restype = None
        lib.fnGetSIMDLevel.argtypes = [C.c_void_p]
        lib.fnGetSIMDLevel.restype = C.c_uint
        node = lib.fnNewFromEncodedNodeTree(NODE_TREE, 0)
        if not node:
            raise RuntimeError("FastNoise2 could not decode the reference node tree")
        lib.fnDeleteNodeRef.argtypes = [C.c_void_p]
        lib.fnDeleteNodeRef.restype = None
        import atexit

        atexit.register(lambda: lib.fnDeleteNodeRef(node))
        _fn = (lib, node)
    return _fn


def generate_sector_fastnoise(sx, sy, sz):
    """TerrainGenerator::GenerateSector, TerrainGenerator.cpp:5-34."""
    lib, node = _fastnoise()
    noise = np.empty(32 * 32 * 32, np.float32)
    # :10 GenUniformGrid3D(buf, 32x, 32y-96, 32z, 32,32,32, 0.004, 12345); output index x + 32y + 1024z
    lib.fnGenUniformGrid3D(node, noise.ctypes.data, sx * 32, sy * 32 - 96, sz * 32, 32, 32, 32, 0.004, 12345, None)
    n_zyx = noise.reshape(32, 32, 32)
    # :21-23 fill = noise < 0 ; id = 245 + (trunc2i(noise * 1234.5678f) & 3)
    grass = 245 + (np.trunc(n_zyx * np.float32(1234.5678)).astype(np.int32) & 3)
    ids = np.where(n_zyx < 0, grass, 0).astype(np.uint8)
    return sector_to_bricks(ids.transpose(1, 0, 2))  # -> [y,z,x]


def _scene_cache_path(tag):
    return CACHE_DIR / f"{tag}.npz"


def save_scene(scene, path):
    keys = sorted(scene["sectors"].keys())
    pos = np.array(keys, np.int32).reshape(-1, 3)
    masks = np.array([scene["sectors"][k][0] for k in keys], np.uint64)
    bricks = np.concatenate([scene["sectors"][k][1] for k in keys], axis=0) if keys else np.zeros((0, 512), np.uint8)
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_name(f".{path.stem}.{os.getpid()}.tmp.npz")
    np.savez(tmp, pos=pos, masks=masks, bricks=bricks, palette=scene["palette"], name=scene.get("name", ""))
    os.replace(tmp, path)  # atomic: a killed writer never leaves a half file behind


def load_scene(path):
    with np.load(path, allow_pickle=False) as z:  # NpzFile re-reads an array on every access
        pos, masks, bricks, palette, name = z["pos"], z["masks"], z["bricks"], z["palette"], str(z["name"])
    sectors = {}
    off = 0
    for p, m in zip(pos, masks):
        k = bin(int(m)).count("1")
        sectors[(int(p[0]), int(p[1]), int(p[2]))] = (int(m), bricks[off : off + k])
        off += k
    return {"sectors": sectors, "palette": palette, "name": name}


def terrain_fastnoise(nx=24, ny=7, nz=24, cache=True):
    """The reference app's start-up terrain (Main.cpp:63-69). ~3 s of FastNoise2 on one core."""
    tag = f"fastnoise_{nx}x{ny}x{nz}"
    cp = _scene_cache_path(tag)
    if cache and cp.exists():
        return load_scene(cp)
    sectors = {}
    coords = [(x, y, z) for y in range(ny) for z in range(nz) for x in range(nx)]
    if len(coords) > 8192:
        # big extents (BASELINE configs[3]): FastNoise2 generation is thread-safe and releases the GIL inside ctypes
        from concurrent.futures import ThreadPoolExecutor

        _fastnoise()
        with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as pool:
            results = pool.map(lambda c: generate_sector_fastnoise(*c), coords, chunksize=64)
            for c, (mask, bricks) in zip(coords, results):
                if mask:
                    sectors[c] = (mask, bricks)
    else:
        for c in coords:
            mask, bricks = generate_sector_fastnoise(*c)
            if mask:
                sectors[c] = (mask, bricks)
    scene = {"sectors": sectors, "palette": reference_palette(), "name": f"FastNoise2 terrain {nx}x{ny}x{nz} sectors"}
    if cache:
        try:
            save_scene(scene, cp)
        except OSError:
            pass
    return scene


def terrain_fastnoise_big(nx=128, ny=24, nz=128, y_shift=512, shared_dir=None, is_writer=True, wait=None):
    """BASELINE configs[3]: the reference's terrain over nx x ny x nz sectors, moved up by y_shift voxels (everything below the surface is
    solid, so the shift sets the byte count: 128 x 24 x 128 sectors with y_shift = 512 hold ~20 M bricks = 10 GB of voxels).  Generated by
    scenes/terrain_gen.c (OpenMP over sectors, FastNoise2 from the reference's vendored tree) into ONE array with a 64-brick block per sector.
    shared_dir (e.g. /dev/shm): the writer process creates the arrays there as files and every other process of the node maps them
    read-only (`wait()` must return once the writer is done) — eight ranks then share one copy of the 13 GB block array."""
    lib, node = _fastnoise()
    gen = C.CDLL(str(HERE / "_ref" / "libterrain_gen.so"))
    gen.terrain_gen_sectors.restype = C.c_uint64
    gen.terrain_gen_sectors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    n = nx * ny * nz
    tag = f"vrt_terrain_{nx}x{ny}x{nz}_s{y_shift}"
    if shared_dir is not None:
        mpath, bpath = Path(shared_dir) / f"{tag}.masks", Path(shared_dir) / f"{tag}.bricks"
        if is_writer:
            masks = np.lib.format.open_memmap(mpath, mode="w+", dtype=np.uint64, shape=(n,))
            bricks = np.lib.format.open_memmap(bpath, mode="w+", dtype=np.uint8, shape=(n * 64, 512))
        else:
            wait()
            masks = np.load(mpath, mmap_mode="r")
            bricks = np.load(bpath, mmap_mode="r")
    else:
        masks = np.zeros(n, np.uint64)
        bricks = np.empty((n * 64, 512), np.uint8)  # (pages of never-written blocks are never touched)
    if shared_dir is None or is_writer:
        fn_ptr = C.cast(lib.fnGenUniformGrid3D, C.c_void_p)
        gen.terrain_gen_sectors(fn_ptr, node, nx, ny, nz, int(y_shift), masks.ctypes.data, bricks.ctypes.data, 0)
        if shared_dir is not None:
            masks.flush()
            bricks.flush()
            if wait is not None:
                wait()
    sectors = {}
    nz_idx = np.nonzero(masks)[0]
    pop = np.array([bin(int(m)).count("1") for m in masks[nz_idx]], np.int64)
    for i, k in zip(nz_idx.tolist(), pop.tolist()):
        x, z, y = i % nx, (i // nx) % nz, i // (nx * nz)
        sectors[(x, y, z)] = (int(masks[i]), bricks[i * 64 : i * 64 + k])
    return {"sectors": sectors, "palette": reference_palette(), "name": f"FastNoise2 terrain {nx}x{ny}x{nz} sectors raised by {y_shift} voxels"}


# ---------------------------------------------------------------------------------------------
# integer-hash terrain: reproducible bit for bit on any host
# ---------------------------------------------------------------------------------------------
def _hash32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _lattice(ix, iy, iz, seed):
    h = _hash32(ix.astype(np.uint32) * np.uint32(0x9E3779B1) ^ _hash32(iy.astype(np.uint32) * np.uint32(0x85EBCA77) ^ _hash32(iz.astype(np.uint32) * np.uint32(0xC2B2AE3D) ^ np.uint32(seed))))
    return (h >> np.uint32(16)).astype(np.int64)  # 0..65535


def _value_noise_fixed(x, y, z, shift, seed):
    """trilinear value noise in 16.16 fixed point; integer arithmetic only."""
    cell = 1 << shift
    ix, iy, iz = x >> shift, y >> shift, z >> shift
    fx, fy, fz = (x & (cell - 1)).astype(np.int64), (y & (cell - 1)).astype(np.int64), (z & (cell - 1)).astype(np.int64)
    acc = np.zeros(np.broadcast(x, y, z).shape, np.int64)
    for dz in (0, 1):
        wz = fz if dz else cell - fz
        for dy in (0, 1):
            wy = fy if dy else cell - fy
            for dx in (0, 1):
                wx = fx if dx else cell - fx
                acc = acc + _lattice(ix + dx, iy + dy, iz + dz, seed) * wx * wy * wz
    return acc >> (3 * shift)  # 0..65535


def terrain_hash(nx=6, ny=4, nz=6, seed=1234, emissive=True):
    """Hilly terrain with overhangs, floating islands and a few emissive blocks.  Pure integer math."""
    sectors = {}
    X = np.arange(nx * 32, dtype=np.int64)
    Z = np.arange(nz * 32, dtype=np.int64)
    xx, zz = np.meshgrid(X, Z, indexing="xy")  # [z,x]
    h = (
        _value_noise_fixed(xx, 0 * xx, zz, 6, seed) * 3 // 4 + _value_noise_fixed(xx, 0 * xx, zz, 4, seed + 1) // 4 + _value_noise_fixed(xx, 0 * xx, zz, 2, seed + 2) // 16
    )  # 0..~70k
    height = (h * (ny * 32 * 5 // 8)) >> 16  # [z,x] ground height in voxels
    for sy in range(ny):
        Y = np.arange(sy * 32, sy * 32 + 32, dtype=np.int64)
        for sz in range(nz):
            for sx in range(nx):
                hh = height[sz * 32 : sz * 32 + 32, sx * 32 : sx * 32 + 32]  # [z,x]
                yy = Y[:, None, None]  # [y,1,1]
                x3 = np.broadcast_to(X[sx * 32 : sx * 32 + 32][None, None, :], (32, 32, 32))
                z3 = np.broadcast_to(Z[sz * 32 : sz * 32 + 32][None, :, None], (32, 32, 32))
                y3 = np.broadcast_to(yy, (32, 32, 32))
                cave = _value_noise_fixed(x3, y3, z3, 4, seed + 7)
                solid = (y3 <= hh[None, :, :]) & (cave > 21000)
                island = (_value_noise_fixed(x3, y3, z3, 3, seed + 11) > 52000) & (y3 > hh[None, :, :] + 12)
                solid = solid | island
                if not solid.any():
                    continue
                hv = _hash32((x3 * 73856093 ^ y3 * 19349663 ^ z3 * 83492791).astype(np.uint32) + np.uint32(seed))
                ids = (1 + (hv % np.uint32(244))).astype(np.uint8)
                ids = np.where(y3 + 2 >= hh[None, :, :], (245 + (hv & np.uint32(3))).astype(np.uint8), ids)
                if emissive:
                    ids = np.where((hv >> np.uint32(8)) % np.uint32(997) == 0, (252 + ((hv >> np.uint32(20)) & np.uint32(3))).astype(np.uint8), ids)
                ids = np.where(solid, ids, 0).astype(np.uint8)
                mask, bricks = sector_to_bricks(ids)
                if mask:
                    sectors[(sx, sy, sz)] = (mask, bricks)
    return {"sectors": sectors, "palette": reference_palette(), "name": f"hash terrain {nx}x{ny}x{nz} sectors seed {seed}"}


def bench_terrain(prefer_fastnoise=True):
    """Scene of BASELINE.json configs 1/2.  Falls back to the hash terrain over the same 24x7x24
    sector range when the FastNoise2 library is not in the tree; says which in scene['name']."""
    if prefer_fastnoise and fastnoise_available():
        try:
            return terrain_fastnoise()
        except Exception as e:  # pragma: no cover
            print(f"[scenes] FastNoise2 terrain failed ({e}); using hash terrain")
    return terrain_hash(24, 7, 24, seed=12345, emissive=False)


def scene_records(scene, dirty_all=True):
    """-> list of (sx,sy,sz,alloc,dirty,bricks) for Context.sync / OracleMap.sync."""
    recs = []
    for (sx, sy, sz), (mask, bricks) in sorted(scene["sectors"].items()):
        recs.append((sx, sy, sz, mask, mask if dirty_all else 0, bricks if dirty_all else None))
    return recs


def scene_stats(scene):
    nb = sum(bin(m).count("1") for m, _ in scene["sectors"].values())
    solid = sum(int(np.count_nonzero(b)) for _, b in scene["sectors"].values()) if nb <= 4_000_000 else None  # (10 GB scenes: not worth a pass)
    return {"sectors": len(scene["sectors"]), "bricks": nb, "solid_voxels": solid, "voxel_bytes": nb * 512, "cell_mask_bytes": nb * 64}


def scene_digest(scene):
    h = hashlib.sha256()
    for k in sorted(scene["sectors"].keys()):
        m, b = scene["sectors"][k]
        h.update(np.array(k, np.int32).tobytes())
        h.update(int(m).to_bytes(8, "little"))
        h.update(np.ascontiguousarray(b).tobytes())
    return h.hexdigest()
