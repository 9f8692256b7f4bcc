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
log = open(out_dir / ("post_check_perf.log" if "--perf-only" in sys.argv else "post_check.log"), "w")


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
PERF_ONLY = "--perf-only" in sys.argv  # under ncu: only the 4K frames
for passes, frames, moving in (() if PERF_ONLY else ((5, 5, True), (0, 4, True), (2, 4, True), (1, 3, False))):
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
# real traced frames (1 bounce, blue noise, sky) from a moving camera; then the one-call trace->window entry point
from scenes import camera, shading, terrain  # noqa: E402
from voxelrt_b200 import capi  # noqa: E402

if not PERF_ONLY:
    scene = terrain.terrain_hash(6, 4, 6, seed=77)
    ctx = capi.Context(6, 4, device=0)
    ctx.set_palette(scene["palette"])
    ctx.sync(terrain.scene_records(scene))
    ctx.set_blue_noise(shading.load_blue_noise()[0])
    sd, st, _ = shading.load_sky()
    ctx.set_sky(sd, st)
    w, h = 256, 144
    gb, gb2, orc = post.GBuffer(0), post.GBuffer(0), pp.PostOracle(w, h)
    worst = {"rgba": 0, "irr": 0, "hist": 0, "render_present": 0}
    for f in range(5):
        cam = camera.Camera(pos=(96.3 + 0.6 * f, 90.2 + 0.1 * f, 20.7 + 0.4 * f), yaw=0.2 + 0.01 * f, pitch=-0.45)
        proj, inv, wo, frac = cam.matrices(w, h)
        fr = capi.make_frame(w, h, inv, proj, wo, frac, frame_no=f + 1, bounces=1)
        out, _ = ctx.render(fr)
        tiles = np.frombuffer(out.tobytes(), dtype=capi.TILE_DTYPE).copy()
        gc = post.make_camera(w, h, proj, inv, cam.pos)
        gb.set_camera(gc)
        orc.set_camera(proj, inv, cam.pos)
        ig, io = gb.denoise_present(tiles), orc.denoise_present(tiles)
        gb2.set_camera(gc)
        ip = gb2.render_present(ctx, capi.make_frame(w, h, inv, proj, wo, frac, frame_no=f + 1, bounces=1))
        d = {"rgba": int((ig != io).sum()), "irr": int((gb.read(0)["irr"] != orc.read(orc.IRR)).any(axis=1).sum()),
             "hist": int((gb.read(4) != orc.read(orc.HIST)).sum()), "render_present": int((ip != io).sum())}
        say(f"traced frame={f} mismatching pixels: {d}; history>0: {float((orc.read(orc.HIST) > 0).mean()):.3f}")
        for k, v in d.items():
            worst[k] = max(worst[k], v)
    res["traced"] = worst
    gb.close()
    gb2.close()
    ctx.close()
    import subprocess  # noqa: E402

    r = subprocess.run([str(ROOT / "tests" / "native" / "test_host"), "--gpu-present"], capture_output=True, text=True, timeout=120)
    say("native:", r.stdout.strip().replace("\n", " | "), r.stderr.strip()[:300])
    res["native_present_ok"] = (r.returncode == 0)
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
for passes in ((5,) if PERF_ONLY else (5, 0)):
    gb = post.GBuffer(0)
    gb.set_passes(passes)
    times = []
    for it in range(6 if PERF_ONLY else 12):
        proj, inv, pos, tiles = seq[it & 1]
        assert rt.cudaMemcpy(d_tiles, tiles.ctypes.data, w * h * 16, 1) == 0
        gb.set_camera(post.make_camera(w, h, proj, inv, pos))
        rt.cudaDeviceSynchronize()
        t = time.perf_counter()
        gb.denoise_present_device(d_tiles.value, d_out.value, 0)
        rt.cudaDeviceSynchronize()
        times.append((time.perf_counter() - t) * 1e3)
    say(f"4K passes={passes}: ms per frame (wall, synced) {[round(x, 3) for x in times]}")
    res[f"ms_4k_passes{passes}"] = float(np.median(times[-4:]))
    gb.close()
say(json.dumps(res))
if not PERF_ONLY:
    (out_dir / "post_check.json").write_text(json.dumps(res))
