#!/usr/bin/env python
"""Kernel-only A/B timing of the 4K primary frame under different vrt_set_option settings (GPU box).

    python tools/exp.py "persistent=0" "persistent=1" "persistent=2,macro_steps=1" ...
Prints ms/frame (CUDA events on the launching stream, L2 flushed between frames, median and min of N).
"""
import os
import sys

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests")
import numpy as np
import torch



from voxelrt_b200 import capi

if os.environ.get('VRT_LIB'):
    capi._lib = capi.load(os.environ['VRT_LIB'])
import bench
from conftest import ctx_for
from scenes import terrain

N = int(os.environ.get("VRT_EXP_N", "40"))
workload = "terrain"
bounces = None
extra_flags = 0
argv = sys.argv[1:]
while argv and argv[0].startswith("--"):
    if argv[0] == "--workload":
        workload = argv[1]
    elif argv[0] == "--bounces":
        bounces = int(argv[1])
    elif argv[0] == "--flags":  # extra VrtFrame.flags, e.g. 16 = VRT_FRAME_GLSL
        extra_flags = int(argv[1])
    argv = argv[2:]
scene, recs, sstats = bench.build_scene(workload)
cap = 1 << 18
while cap < sstats["bricks"] + 4096:
    cap <<= 1
ctx = capi.Context(*bench._WL["view"], device=0, initial_brick_capacity=cap)
ctx.set_palette(scene["palette"])
ctx.sync(recs)
_, w, h, b0 = bench.WORKLOADS[workload]
bounces = b0 if bounces is None else bounces
if bounces:
    from scenes import shading

    ctx.set_blue_noise(shading.load_blue_noise()[0])
    d_, t_, _ = shading.load_sky()
    ctx.set_sky(d_, t_)
fb = torch.zeros(w * h * 4, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
frame = bench.bench_frame(w, h, bounces)
frame.flags |= extra_flags
rays_per_px = (1 + bounces) if not (extra_flags & 16) else (1 + (1 if bounces else 0) + bounces + min(bounces, 2))
sets = argv or ["persistent=0"]
ref = None
for rep in range(2):
    for spec in sets:
        for kv in spec.split(","):
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
        for _ in range(5):
            ctx.render_device(frame, fb.data_ptr(), None, st.cuda_stream)
        ts = []
        for _ in range(N):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            ctx.render_device(frame, fb.data_ptr(), None, st.cuda_stream)
            b.record(st)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        digest = int(fb.to(torch.int64).sum().item())
        if ref is None:
            ref = digest
        ts = np.array(ts)
        print(f"{workload} b{bounces} {spec:34s} median {np.median(ts):.4f} ms  min {ts.min():.4f} ms  -> {w*h*rays_per_px/np.median(ts)/1e6:.2f} Grays/s (nominal rays)  same_frame={digest == ref}", flush=True)
