import sys, time
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from conftest import ctx_for
from scenes import camera, terrain
from voxelrt_b200 import capi
import bench
scene = terrain.bench_terrain()
ctx = ctx_for(scene)
w, h = 3840, 2160
fb = torch.zeros(w * h * 4, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
frame = bench.bench_frame(w, h, 0)
for mode in (0, 1, 2):
    ctx.set_option("macro_steps", mode)
    ctx.set_option("metrics", 1)
    ctx.render_device(frame, fb.data_ptr(), None, st.cuda_stream); torch.cuda.synchronize()
    m = ctx.metrics()
    if mode == 2:
        import ctypes as C
        arr = (C.c_uint64 * 8)()
        ctx.lib.vrt_debug_macro_diag(ctx.h, arr)
        names = ["attempts", "fail_t1<=tcur", "fail_tau", "fail_side", "fail_chord", "jumps", "sum_man", "ambig"]
        print("   diag", {n: int(v) for n, v in zip(names, arr)})
    print("macro", mode, "rays", m.rays, "iters", m.iters, "sector(or tries)", m.sector_fetches, "cell(or jumps)", m.cell_fetches, "hits", m.hits, "capped", m.capped)
    ctx.set_option("metrics", 0)
    for _ in range(3): ctx.render_device(frame, fb.data_ptr(), None, st.cuda_stream)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(10): ctx.render_device(frame, fb.data_ptr(), None, st.cuda_stream)
    b.record(st); torch.cuda.synchronize()
    print("   ms/frame", a.elapsed_time(b) / 10)
# box statistics
import ctypes as C
n_has = 0; ext = []
for (sx, sy, sz) in [(16,4,16),(16,8,16),(5,10,40),(30,12,30),(0,15,0),(16,5,16),(20,6,20)]:
    mask, base, _, _ = ctx.read_sector(sx, sy, sz)
    print("sector", (sx,sy,sz), "mask", hex(mask), "z(base/boxlo)", hex(base))
