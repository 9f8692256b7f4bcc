#!/usr/bin/env python
"""Kernel-only A/B timing of the 4K primary frame under different vrt_set_option settings (GPU box).

    python tools_exp.py "persistent=0" "persistent=1" "persistent=2,macro_steps=1" ...
Prints ms/frame (CUDA events on the launching stream, L2 flushed between frames, median and min of N).
"""
import sys

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests")
import numpy as np
import torch

import os

from voxelrt_b200 import capi

if os.environ.get('VRT_LIB'):
    capi._lib = capi.load(os.environ['VRT_LIB'])
import bench
from conftest import ctx_for
from scenes import terrain

N = 40
scene = terrain.bench_terrain()
ctx = ctx_for(scene, initial_brick_capacity=1 << 18)
w, h = 3840, 2160
fb = torch.zeros(w * h * 4, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
frame = bench.bench_frame(w, h, 0)
sets = sys.argv[1:] or ["persistent=1"]
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
        print(f"{spec:40s} median {np.median(ts):.4f} ms  min {ts.min():.4f} ms  -> {w*h/np.median(ts)/1e6:.2f} Grays/s  same_frame={digest == ref}", flush=True)
