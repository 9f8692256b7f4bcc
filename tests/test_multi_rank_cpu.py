"""The N>1 path on CPU: two processes over torch.distributed/gloo each take their share of the screen
tiles (voxelrt_b200.partition), render it with the oracle's tile-split mode (same rule as the CUDA
kernel), and the gathered frame must equal the single-rank frame byte for byte."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, rows=False):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from oracle import pyoracle
    from scenes import camera, terrain
    from voxelrt_b200 import capi, partition

    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = terrain.terrain_hash(4, 3, 4, seed=5)
    orc = pyoracle.OracleMap(6, 4)
    orc.set_palette(scene["palette"])
    orc.sync(terrain.scene_records(scene))
    w, h = 200, 136  # not a multiple of 32: clipped tiles on both edges
    cam = camera.Camera(pos=(60.2, 70.1, 10.3), yaw=0.3, pitch=-0.5)
    proj, inv, wo, frac = cam.matrices(w, h)
    flags = capi.VRT_FRAME_LINEAR_OUTPUT | (capi.VRT_FRAME_PART_ROWS if rows else 0)
    part, _, st = orc.render(capi.make_frame(w, h, inv, proj, wo, frac, flags=flags, part_index=rank, part_count=world), threads=2)
    # every rank's ray count is what the host-side partition predicts
    assert st.rays == partition.pixels_of_rank(w, h, rank, world, rows)
    mine = torch.tensor([r[1] * w + r[0] for r in partition.rects_of_rank(w, h, rank, world, rows)], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()]))
    # gather = sum of the disjoint partial frames (unowned pixels are zero)
    t = torch.from_numpy(part.astype(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    owned = torch.from_numpy((part != 0).any(axis=0).astype(np.int64))
    dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    if rank == 0:
        full, _, st_full = orc.render(capi.make_frame(w, h, inv, proj, wo, frac, flags=capi.VRT_FRAME_LINEAR_OUTPUT), threads=2)
        tx, ty = partition.tile_grid(w, h)
        ok = (
            int(sum(int(s) for s in sizes)) == ((h + partition.BAND - 1) // partition.BAND if rows else tx * ty)
            and np.array_equal(t.numpy().astype(np.uint32), full)
            and int(owned.max()) <= 1
            and st_full.rays == w * h
        )
        Path(out_dir, "result.txt").write_text("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,rows", [(2, False), (3, False), (2, True), (3, True)])
def test_tile_split_gather_gloo(tmp_path, world, rows):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), rows), nprocs=world, join=True)
    assert (tmp_path / "result.txt").read_text() == "ok"


def test_partition_covers_every_pixel_once():
    from voxelrt_b200 import partition

    for w, h, n in [(3840, 2160, 8), (1920, 1080, 4), (36, 4, 2), (100, 100, 7)]:
        seen = np.zeros((h, w), np.int32)
        for r in range(n):
            for t in partition.tiles_of_rank(w, h, r, n):
                x0, y0, tw, th = partition.tile_rect(w, h, t)
                seen[y0 : y0 + th, x0 : x0 + tw] += 1
            assert partition.pixels_of_rank(w, h, r, n) == sum(partition.tile_rect(w, h, t)[2] * partition.tile_rect(w, h, t)[3] for t in partition.tiles_of_rank(w, h, r, n))
        assert (seen == 1).all()
        # row-band split: bands are disjoint, cover the frame, and each is one contiguous byte range of the tile layout
        seen[:] = 0
        for r in range(n):
            for x0, y0, tw, th in partition.rects_of_rank(w, h, r, n, rows=True):
                assert (y0 // partition.BAND) % n == r and x0 == 0 and tw == w
                seen[y0 : y0 + th, x0 : x0 + tw] += 1
        assert (seen == 1).all()
        assert sum(partition.pixels_of_rank(w, h, r, n, rows=True) for r in range(n)) == w * h
