"""The reference's voxel-map cache file, "cvox 0004" (VoxelRT/VoxelMap.cpp:205-274 on LibGlimpsw/Common/BinaryIO.cpp:8-65) —
read and written wire-compatibly, so big scenes (BASELINE configs[3] takes minutes of FastNoise2) can be cached in the file the
reference itself loads at start-up (`logs/voxels_2k_sponza.dat`, Main.cpp:38-49).  INPUT/OUTPUT format code on either side of
the hot path; not the product path, not the oracle.

Layout (little endian):
    u64  magic  0x00000004'786f7663  ("cvox", 4)
    u32  number of sectors
    blob Palette: Material[256], 16 bytes each {u8 r, g, b, fuzz; f32 emission; 8 bytes padding}      (VoxelMap.h:22-26)
    packs until every sector has been read; a pack is  u32 raw size + blob  and holds whole sector records
         u32 sector index = x & 0xFFF | (z & 0xFFF) << 12 | (y & 0xFF) << 24   (WorldSectorIndexer, VoxelMap.h:60-100)
         u64 allocation mask
         512 bytes per set bit, ascending                                       (packs are cut after >= 16 MiB)
    blob = u32 compressed size + one zstd frame (default level, content checksum on)

zstd comes from the system's libzstd.so.1 through ctypes (its development header is absent from this image).
`tests/test_scenes_cpu.py` pins both directions against the reference's own Serialize / Deserialize (oracle/_ref).
"""
from __future__ import annotations

import ctypes as C
import struct
from pathlib import Path

import numpy as np

MAGIC = 0x00000004786F7663
MAX_PACK = 16 << 20
_z = None


def _zstd():
    global _z
    if _z is None:
        lib = C.CDLL("libzstd.so.1")
        lib.ZSTD_compressBound.argtypes = [C.c_size_t]
        lib.ZSTD_compressBound.restype = C.c_size_t
        lib.ZSTD_createCCtx.restype = C.c_void_p
        lib.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        lib.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.ZSTD_CCtx_setParameter.restype = C.c_size_t
        lib.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.ZSTD_compress2.restype = C.c_size_t
        lib.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.ZSTD_decompress.restype = C.c_size_t
        lib.ZSTD_isError.argtypes = [C.c_size_t]
        lib.ZSTD_isError.restype = C.c_uint
        _z = lib
    return _z


def _compress(raw: bytes) -> bytes:
    z = _zstd()
    ctx = z.ZSTD_createCCtx()
    try:
        z.ZSTD_CCtx_setParameter(ctx, 100, 3)  # ZSTD_c_compressionLevel = ZSTD_CLEVEL_DEFAULT (BinaryIO.cpp:10)
        z.ZSTD_CCtx_setParameter(ctx, 201, 1)  # ZSTD_c_checksumFlag (BinaryIO.cpp:11)
        cap = z.ZSTD_compressBound(len(raw))
        out = C.create_string_buffer(cap)
        n = z.ZSTD_compress2(ctx, out, cap, raw, len(raw))
        if z.ZSTD_isError(n):
            raise IOError("zstd compression failed")
        return out.raw[:n]
    finally:
        z.ZSTD_freeCCtx(ctx)


def _decompress(blob: bytes, raw_size: int) -> bytes:
    z = _zstd()
    out = C.create_string_buffer(max(raw_size, 1))
    n = z.ZSTD_decompress(out, raw_size, blob, len(blob))
    if z.ZSTD_isError(n):
        raise IOError("Failed to decompress stream")
    if n != raw_size:
        raise IOError("Decompressed stream is too short")  # BinaryIO.cpp:61-63
    return out.raw[:raw_size]


def sector_index(x, y, z):
    return (x & 0xFFF) | ((z & 0xFFF) << 12) | ((y & 0xFF) << 24)


def sector_pos(idx):
    def sx(v, bits):
        return v - (1 << bits) if v >> (bits - 1) else v

    return sx(idx & 0xFFF, 12), sx((idx >> 24) & 0xFF, 8), sx((idx >> 12) & 0xFFF, 12)


def decode_palette(enc: np.ndarray):
    """encoded u64[256] (Material::GetEncoded) -> Material records (r, g, b, fuzz, emission); RGB565 expands by shifting, which
    GetEncoded maps back onto the same bits."""
    out = []
    for e in np.asarray(enc, np.uint64).tolist():
        r, g, b = ((e >> 11) & 31) << 3, ((e >> 5) & 63) << 2, (e & 31) << 3
        emission = float(np.array([(e >> 16) & 0xFFFF], np.uint16).view(np.float16)[0])
        out.append((r, g, b, (e >> 32) & 0xFF, emission))
    return out


def save_cvox(scene, path):
    """scene: {"sectors": {(sx, sy, sz): (alloc_mask, bricks[k, 512])}, "palette": u64[256] encoded} -> cvox 0004 file."""
    mats = scene.get("materials") or decode_palette(scene["palette"])
    pal = b"".join(struct.pack("<BBBBf8x", r, g, b, f, e) for r, g, b, f, e in mats)
    assert len(pal) == 4096
    with open(path, "wb") as f:
        f.write(struct.pack("<QI", MAGIC, len(scene["sectors"])))
        blob = _compress(pal)
        f.write(struct.pack("<I", len(blob)) + blob)
        pack = bytearray()

        def flush(final=False):
            nonlocal pack
            if len(pack) >= MAX_PACK or final:
                blob = _compress(bytes(pack))
                f.write(struct.pack("<II", len(pack), len(blob)) + blob)
                pack = bytearray()

        for (sx, sy, sz), (mask, bricks) in scene["sectors"].items():
            pack += struct.pack("<IQ", sector_index(sx, sy, sz), mask)
            pack += np.ascontiguousarray(bricks, np.uint8).tobytes()
            flush()
        flush(final=True)  # VoxelMap.cpp:273 always writes a last pack, empty or not


def load_cvox(path):
    """cvox 0004 file -> scene dict (plus "materials": the raw Material records)."""
    from . import terrain

    data = Path(path).read_bytes()
    if len(data) < 12:
        raise IOError("End of stream")
    magic, n_sectors = struct.unpack_from("<QI", data, 0)
    if magic != MAGIC:
        raise IOError("Incompatible file")  # VoxelMap.cpp:219
    p = 12
    (csz,) = struct.unpack_from("<I", data, p)
    pal = _decompress(data[p + 4 : p + 4 + csz], 4096)
    p += 4 + csz
    mats = [struct.unpack_from("<BBBBf", pal, 16 * i) for i in range(256)]
    sectors = {}
    pack, q = b"", 0
    for _ in range(n_sectors):
        if q >= len(pack):
            raw, csz = struct.unpack_from("<II", data, p)
            pack, q = _decompress(data[p + 8 : p + 8 + csz], raw), 0
            p += 8 + csz
        idx, mask = struct.unpack_from("<IQ", pack, q)
        q += 12
        k = bin(mask).count("1")
        bricks = np.frombuffer(pack, np.uint8, count=k * 512, offset=q).reshape(k, 512).copy()
        q += k * 512
        sectors[sector_pos(idx)] = (mask, bricks)
    palette = np.array([terrain.encode_material(r, g, b, fz, em) for r, g, b, fz, em in mats], np.uint64)
    return {"sectors": sectors, "palette": palette, "materials": mats, "name": f"cvox {Path(path).name}"}
