#!/usr/bin/env python
"""GPU-box check: every way of rendering one big two-bounce frame gives the same bytes (device / host-buffer path, one thread per pixel /
wavefront, whole frame / 8-part band split assembled in one host buffer)."""
import ctypes as C
import sys

sys.path.insert(0, "/root/repo")
import numpy as np
import torch

import bench
from voxelrt_b200 import capi

scene, recs, sstats = bench.build_scene("large", large=(128, 7, 128, 0))
ctx = capi.Context(*bench._WL["view"], device=0, initial_brick_capacity=1 << 22)
ctx.set_palette(scene["palette"])
ctx.sync(recs)
from scenes import shading

ctx.set_blue_noise(shading.load_blue_noise()[0])
d_, t_, _ = shading.load_sky()
ctx.set_sky(d_, t_)
w, h, b = 3840, 2160, 2
npx = w * h
st = torch.cuda.Stream()
res = {}
for wave in (0, 1):
    ctx.set_option("wavefront", wave)
    fb = torch.zeros(npx * 4, dtype=torch.int32, device="cuda")
    ctx.render_device(bench.bench_frame(w, h, b), fb.data_ptr(), None, st.cuda_stream)
    torch.cuda.synchronize()
    res[f"device wave={wave}"] = fb.cpu().numpy().copy()
    out, _ = ctx.render(bench.bench_frame(w, h, b))
    res[f"host wave={wave}"] = np.frombuffer(out.tobytes(), np.int32).copy()
    host = np.zeros(npx * 4, np.int32)
    for p in range(8):
        f = bench.bench_frame(w, h, b, part_index=p, part_count=8, flags=capi.VRT_FRAME_PART_ROWS)
        ctx._chk(ctx.lib.vrt_render(ctx.h, C.byref(f), host.ctypes.data, None))
    res[f"host 8 parts wave={wave}"] = host
ctx.set_option("wavefront", 2)
host = np.zeros(npx * 4, np.int32)
for rep in range(4):  # self-tuning in flight
    for p in range(8):
        f = bench.bench_frame(w, h, b, part_index=p, part_count=8, flags=capi.VRT_FRAME_PART_ROWS)
        ctx._chk(ctx.lib.vrt_render(ctx.h, C.byref(f), host.ctypes.data, None))
    res[f"host 8 parts self-tuning rep {rep}"] = host.copy()
    out, _ = ctx.render(bench.bench_frame(w, h, b))
    res[f"host whole self-tuning rep {rep}"] = np.frombuffer(out.tobytes(), np.int32).copy()
ref = res["device wave=0"]
for k, v in res.items():
    bad = np.nonzero(v != ref)[0]
    print(f"{k:40s} differing words: {bad.size}" + (f"  first {bad[:6]} tiles {np.unique(bad[:2000] // 64)[:8]}" if bad.size else ""))
