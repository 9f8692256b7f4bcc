"""Quick GPU check of the GBuffer step without importing torch (a fresh box pages torch in for a minute): parity against the
image-space oracle on a few sequences, then wall-clock timing of 4K frames on device buffers.  Output: gpurun_out/post_check.log"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
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
    # vrt_trace_glsl (the GLSL renderer's rayCast / rayCastCoarse) against orc_trace_glsl
    from oracle import pyoracle  # noqa: E402

    omap = pyoracle.OracleMap(6, 4)
    omap.set_palette(scene["palette"])
    omap.sync(terrain.scene_records(scene))

    def hits_differ(a, b):
        n = 0
        for name in a.dtype.names:
            x, y = a[name], b[name]
            if x.dtype.kind == "f":
                nan = np.isnan(x) & np.isnan(y)
                n += int(((x.view(np.uint32) != y.view(np.uint32)) & ~nan).sum())
            else:
                n += int((x != y).sum())
        return n

    rng = np.random.default_rng(11)
    vals = np.array([0.0, -0.0, 1.0, -1.0, 1e-39, -1e-39, 1e30, np.inf, -np.inf, np.nan, 0.3, -0.7], np.float32)
    d_sp = np.array([(a, b, c) for a in vals for b in vals for c in vals], np.float32)
    o_sp = np.tile(np.array([[0.25, 0.5, 0.75]], np.float32), (len(d_sp), 1))
    o_sp[::7, 0] = np.nan
    glsl = {}
    for flags in (0, 1, 2, 3):
        bad = 0
        for k in range(3):
            wo_ = (int(rng.integers(2, 190)), int(rng.integers(2, 126)), int(rng.integers(2, 190)))
            o_ = rng.random((20000, 3)).astype(np.float32)
            d_ = rng.normal(size=(20000, 3))
            d_ = (d_ / np.linalg.norm(d_, axis=1, keepdims=True)).astype(np.float32)
            bad += hits_differ(ctx.trace_glsl(o_, d_, wo_, flags), omap.trace_glsl(o_, d_, wo_, flags)[0])
        o_ = (rng.uniform(-100, 100, (20000, 3))).astype(np.float32)  # far origins (bounce-like) around (96, 64, 96)
        bad += hits_differ(ctx.trace_glsl(o_, d_, (96, 64, 96), flags), omap.trace_glsl(o_, d_, (96, 64, 96), flags)[0])
        o_ = (rng.random((5000, 3)) - 0.5).astype(np.float32)  # camera above the view box
        d2 = rng.normal(size=(5000, 3)) * 0.08
        d2[:, 1] = -1.0
        d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
        want = omap.trace_glsl(o_, d2, (96, 600, 96), flags)[0]
        bad += hits_differ(ctx.trace_glsl(o_, d2, (96, 600, 96), flags), want)
        bad_sp = hits_differ(ctx.trace_glsl(o_sp, d_sp, (90, 70, 90), flags), omap.trace_glsl(o_sp, d_sp, (90, 70, 90), flags)[0])
        glsl[f"flags{flags}"] = {"mismatching_fields": bad, "special_directions": bad_sp, "outside_hit_fraction": float(((want["flags"] & 0x100) != 0).mean())}
        say(f"trace_glsl flags={flags}: {glsl[f'flags{flags}']}")
    recs2 = [(40, 10, 40, 1 << 21, 1 << 21, np.full((1, 512), 250, np.uint8)), (2, 1, 2, 0, 0xFFFFFFFFFFFFFFFF, None, True)]
    ctx.sync(recs2)
    omap.sync(recs2)
    wo2 = (40 * 32 + 5, 10 * 32 + 60, 40 * 32 + 7)
    o2 = rng.random((4000, 3)).astype(np.float32)
    d3 = np.tile(np.array([[0.05, -1.0, 0.08]], np.float32), (4000, 1)) + rng.normal(size=(4000, 3)).astype(np.float32) * 0.1
    want = omap.trace_glsl(o2, d3, wo2, 0)[0]
    glsl["after_edit"] = {"mismatching_fields": hits_differ(ctx.trace_glsl(o2, d3, wo2, 0), want) + hits_differ(ctx.trace_glsl(o_, d2, (96, 64, 96), 1), omap.trace_glsl(o_, d2, (96, 64, 96), 1)[0]),
                          "lone_brick_hits": int(((want["flags"] & 0x100) != 0).sum())}
    say(f"trace_glsl after edits: {glsl['after_edit']}")
    res["trace_glsl"] = glsl
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
