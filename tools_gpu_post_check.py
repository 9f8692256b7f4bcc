"""Quick GPU check of the GBuffer step without importing torch (a fresh box pages torch in for a minute): parity against the
image-space oracle on a few sequences, then wall-clock timing of 4K frames on device buffers.  Output: gpurun_out/post_check.log"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
out_dir = ROOT / "gpurun_out"
out_dir.mkdir(exist_ok=True)
log = open(out_dir / "post_check.log", "w")


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + "\n")
    log.flush()


from oracle import pypostoracle as pp  # noqa: E402
from scenes import gbuffer_synth as pu  # noqa: E402
from voxelrt_b200 import post  # noqa: E402

res = {}
t0 = time.time()
for passes, frames, moving in ((5, 5, True), (0, 4, True), (2, 4, True), (1, 3, False)):
    w, h = 96, 64
    seq = pu.synthetic_sequence(w, h, frames, seed=100 + passes, moving=moving)
    gb, orc = post.GBuffer(0), pp.PostOracle(w, h)
    gb.set_passes(passes)
    orc.set_passes(passes)
    worst = {}
    for f, (proj, inv, pos, tiles) in enumerate(seq):
        gb.set_camera(post.make_camera(w, h, proj, inv, pos, reset_history=(f == 3)))
        orc.set_camera(proj, inv, pos, reset_history=(f == 3))
        ig, io = gb.denoise_present(tiles), orc.denoise_present(tiles)
        d = {"rgba": int((ig != io).sum())}
        for pg, po, name in ((0, orc.IRR, "irr"), (1, orc.PREV_IRR, "prev"), (2, orc.TEMP_IRR, "temp")):
            d[name] = int((gb.read(pg)["irr"] != orc.read(po)).any(axis=1).sum())
        d["moments"] = int((gb.read(3) != orc.read(orc.MOMENTS)).any(axis=1).sum())
        d["hist"] = int((gb.read(4) != orc.read(orc.HIST)).sum())
        say(f"passes={passes} frame={f} mismatching pixels: {d}")
        for k, v in d.items():
            worst[k] = max(worst.get(k, 0), v)
    res[f"passes{passes}"] = worst
    gb.close()
say("parity seconds", round(time.time() - t0, 2))

# timing at 3840x2160 on device buffers (wall clock around K frames with a device synchronise on both sides)
try:
    rt = C.CDLL("libcudart.so")
except OSError:
    rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
w, h = 3840, 2160
seq = pu.synthetic_sequence(w, h, 2, seed=9)
d_tiles, d_out = C.c_void_p(), C.c_void_p()
assert rt.cudaMalloc(C.byref(d_tiles), w * h * 16) == 0 and rt.cudaMalloc(C.byref(d_out), w * h * 4) == 0
for passes in (5, 0):
    gb = post.GBuffer(0)
    gb.set_passes(passes)
    times = []
    for it in range(12):
        proj, inv, pos, tiles = seq[it & 1]
        assert rt.cudaMemcpy(d_tiles, tiles.ctypes.data, w * h * 16, 1) == 0
        gb.set_camera(post.make_camera(w, h, proj, inv, pos))
        rt.cudaDeviceSynchronize()
        t = time.perf_counter()
        gb.denoise_present_device(d_tiles.value, d_out.value, 0)
        rt.cudaDeviceSynchronize()
        times.append((time.perf_counter() - t) * 1e3)
    say(f"4K passes={passes}: ms per frame (wall, synced) {[round(x, 3) for x in times]}")
    res[f"ms_4k_passes{passes}"] = float(np.median(times[4:]))
    gb.close()
say(json.dumps(res))
(out_dir / "post_check.json").write_text(json.dumps(res))
