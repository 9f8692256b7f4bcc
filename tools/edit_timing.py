#!/usr/bin/env python
"""Where an edit frame's time goes (GPU box): host part of vrt_sync, its device work, the frame kernel."""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import bench
from conftest import ctx_for
from scenes import terrain, edits
from voxelrt_b200 import capi

scene = terrain.bench_terrain()
ctx = ctx_for(scene, initial_brick_capacity=1 << 18)
if len(sys.argv) > 2:
    ctx.set_option("gather_threads", int(sys.argv[2]))
    print("gather_threads", sys.argv[2])
frames, _ = edits.random_edit_frames(scene, 40, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, seed=1)
batches = [capi.make_records(r) for r in frames]
w, h = 3840, 2160
fb = torch.zeros(w * h * 4, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
frame = bench.bench_frame(w, h, 0)
host, dev, ren = [], [], []
for i, (arr, keep, n) in enumerate(batches):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.sync_records(arr, n)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ctx.render_device(frame, fb.data_ptr(), None, st.cuda_stream)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if i >= 5:
        host.append(t1 - t0); dev.append(t2 - t1); ren.append(t3 - t2)
s = ctx.stats()
print(f"records/frame {np.mean([b[2] for b in batches]):.0f}  host part of vrt_sync {1e3*np.median(host):.3f} ms  device tail {1e3*np.median(dev):.3f} ms  frame {1e3*np.median(ren):.3f} ms")
# back-to-back (pipelined) loop
torch.cuda.synchronize(); t0 = time.perf_counter()
for arr, keep, n in batches[:30]:
    ctx.sync_records(arr, n)
    ctx.render_device(frame, fb.data_ptr(), None, st.cuda_stream)
torch.cuda.synchronize(); print(f"back-to-back: {1e3*(time.perf_counter()-t0)/30:.3f} ms per edit frame")
# the same loop when this GPU renders only one eighth of the frame (one rank of 8): what bounds the edit workload at N = 8
part = bench.bench_frame(w, h, 0, part_index=0, part_count=8, flags=capi.VRT_FRAME_PART_ROWS)
torch.cuda.synchronize(); t0 = time.perf_counter()
for arr, keep, n in batches[:30]:
    ctx.sync_records(arr, n)
    ctx.render_device(part, fb.data_ptr(), None, st.cuda_stream)
torch.cuda.synchronize(); print(f"back-to-back, 1/8 of the frame: {1e3*(time.perf_counter()-t0)/30:.3f} ms per edit frame")
hs = []
for arr, keep, n in batches[:30]:
    t0 = time.perf_counter()
    ctx.sync_records(arr, n)
    hs.append(time.perf_counter() - t0)
    ctx.render_device(part, fb.data_ptr(), None, st.cuda_stream)
torch.cuda.synchronize()
print(f"host time inside vrt_sync while pipelined: median {1e3*np.median(hs):.3f} ms  min {1e3*np.min(hs):.3f}  max {1e3*np.max(hs):.3f}")
print("stats", {k: getattr(s, k) for k in ("resident_bricks", "free_ranges", "bricks_uploaded", "bricks_relocated", "last_launches")})
ctx.close()
