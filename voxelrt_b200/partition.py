"""Host-side screen-tile partition: macro tile t (32x32 px, row-major) belongs to rank
t % part_count — the rule `warp_tile_origin` in csrc/vrt_kernels.cuh applies on the device and
`orc_render` applies in the oracle.  With `rows=True` (VRT_FRAME_PART_ROWS) the unit is a BAND of 8 pixel
rows (two tile rows, contiguous in the tile-layout framebuffer): band b belongs to rank b % part_count.
Used by bench.py (per-rank ray counts) and the multi-rank tests."""
from __future__ import annotations

TILE = 32
BAND = 8  # VRT_BAND_ROWS


def tile_grid(width: int, height: int):
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def tiles_of_rank(width: int, height: int, part_index: int, part_count: int):
    tx, ty = tile_grid(width, height)
    return list(range(part_index, tx * ty, part_count))


def bands_of_rank(height: int, part_index: int, part_count: int):
    return list(range(part_index, (height + BAND - 1) // BAND, part_count))


def rects_of_rank(width: int, height: int, part_index: int, part_count: int, rows: bool = False):
    """-> [(x0, y0, w, h)] pixel rectangles owned by the rank (macro tiles, or full-width bands with rows=True)."""
    if rows:
        return [(0, b * BAND, width, min(BAND, height - b * BAND)) for b in bands_of_rank(height, part_index, part_count)]
    return [tile_rect(width, height, t) for t in tiles_of_rank(width, height, part_index, part_count)]


def tile_rect(width: int, height: int, t: int):
    """-> (x0, y0, w, h) of macro tile t clipped to the frame."""
    tx, _ = tile_grid(width, height)
    x0, y0 = (t % tx) * TILE, (t // tx) * TILE
    return x0, y0, min(TILE, width - x0), min(TILE, height - y0)


def band_bytes(width: int) -> int:
    """Bytes of one full band of a tile-layout framebuffer (16 B per pixel)."""
    return width * BAND * 16


def pixels_of_rank(width: int, height: int, part_index: int, part_count: int, rows: bool = False) -> int:
    return sum(w * h for _, _, w, h in rects_of_rank(width, height, part_index, part_count, rows))
