#!/usr/bin/env python
"""Find rays whose result differs between macro_steps=1 and macro_steps=0 (GPU box) and dump them for CPU analysis."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import bench
from voxelrt_b200 import capi

workload = sys.argv[1] if len(sys.argv) > 1 else "sponza"
scene, recs, sstats = bench.build_scene(workload)
cap = 1 << 18
while cap < sstats["bricks"] + 4096:
    cap <<= 1
view = bench._WL["view"]
ctx = capi.Context(*view, device=0, initial_brick_capacity=cap)
ctx.set_palette(scene["palette"]); ctx.sync(recs)
rng = np.random.default_rng(5)
ext_xz, ext_y = 32 << view[0], 32 << view[1]
keys = np.array(list(scene["sectors"].keys()))
lo, hi = keys.min(0) * 32, (keys.max(0) + 1) * 32
fails = []
for rep in range(6):
    n = 2_000_000
    wo = np.array([rng.integers(lo[0], hi[0]), rng.integers(lo[1], hi[1]), rng.integers(lo[2], hi[2])])
    p = np.stack([rng.uniform(lo[a] - 40, hi[a] + 40, n) for a in range(3)], 1)
    o = (p - wo).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True); d = d.astype(np.float32)
    ctx.set_option("macro_steps", 0); a = ctx.trace(o, d, wo)
    ctx.set_option("macro_steps", 1); b = ctx.trace(o, d, wo)
    bad = np.zeros(n, bool)
    for name in a.dtype.names:
        x, y = a[name], b[name]
        if name == "flags":
            x, y = x & 0xFFFF, y & 0xFFFF
        if x.dtype.kind == "f":
            bad |= (x.view(np.uint32) != y.view(np.uint32)) & ~(np.isnan(x) & np.isnan(y))
        else:
            bad |= x != y
    idx = np.nonzero(bad)[0]
    print(f"rep {rep}: wo {wo.tolist()} mismatches {idx.size} of {n}", flush=True)
    for i in idx[:20]:
        fails.append((wo.copy(), o[i].copy(), d[i].copy(), a[i:i+1].copy(), b[i:i+1].copy()))
if fails:
    np.savez("gpurun_out/macro_fail.npz", wo=np.array([f[0] for f in fails]), o=np.array([f[1] for f in fails]), d=np.array([f[2] for f in fails]),
             a=np.concatenate([f[3] for f in fails]), b=np.concatenate([f[4] for f in fails]))
    for f in fails[:6]:
        print("wo", f[0].tolist(), "o", f[1].tolist(), "d", f[2].tolist()); print("   stepwise", f[3]); print("   macro   ", f[4])
