#!/usr/bin/env python
"""bench.py — the BASELINE.json metric: Mrays/s of the brickmap traversal path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload at N=1 = BASELINE.json configs[1]: the reference's FastNoise2 terrain (24x7x24 sectors),
reference camera, 3840x2160 primary rays + normals/material/depth G-buffer.  One step = one frame
(one launch of the traversal kernel over all 8,294,400 rays).

value   device-resident throughput: CUDA events around each frame's kernel on the launching stream,
        L2 flushed (256 MiB memset) between frames, max over ranks.
e2e     the same frame through the host-buffer ABI call (vrt_render): frame constants in, the G-buffer
        copied back into page-locked host memory, wall clock around the blocking call.  Primary-only
        frames move the compact 8 B/px payload (VRT_FRAME_COMPACT: the other half of the reference's
        16 B/px tile is the constant 1.0; value_full_16B_per_px times the full tile as well); at N > 1 all
        ranks deliver their bands into ONE host frame in shared page-locked memory.
roofline  bound "issue" (the kernels are instruction-issue bound, ncu numbers beside it); achieved / peak /
        frac = algorithmic bytes (8 I_s + 8 I_c + 9 H + 16 P, counted by the kernel's own traversal
        counters, which tests check against the oracle) / kernel time vs the measured HBM peak, a
        work-equivalent figure; traffic = DRAM bytes of the committed ncu capture (profiles/ncu_summary.json).
cpu_baseline  the CPU path (oracle/_ref = the reference's own CpuRenderer.cpp when it was compiled
        here, built with the flags it ships with, else the oracle port) on all host threads, same frame.
--impl reference  that CPU path alone, as a bench line of its own (rank 0 only under torchrun, all host threads).
N > 1   screen split of the SAME frame (strong scaling): rank r renders the 8-pixel bands b with
        b % N == r of the replicated brickmap into its own buffer, and vrt_render_gather moves them
        into the presenting GPU's framebuffer with one strided copy over NVLink while the next frame
        is traced (presenting GPU = frame % N).  value = K frames back to back including the last
        gather, max over ranks; gather.* carries the non-pipelined frame latency and a trace-only figure.
--workload  sponza | large | edits: the other BASELINE configs (non-default bench lines).
--present   additionally time the step after the path (the reference's GBuffer: blit + reprojection + SVGF + present,
        include/voxelrt_b200_post.h) on the traced frames and add a "present" block (N = 1 only; on by default, --no-present skips it).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_summary(args):
    """What the committed ncu capture of the dominant kernel says (profiles/ncu_summary.json, written by tools/profiles.py from an
    `ncu --set full` capture of the same command): DRAM / L2 bytes per launch, issue-slot and pipe utilisation, executed instructions.
    None for shapes without a capture."""
    try:
        d = json.loads((ROOT / "profiles" / "ncu_summary.json").read_text())
        key = f"{args.workload} {args.width}x{args.height} bounces={args.bounces} gpus={args.gpus}"
        return d.get(key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.samples = []
            self.stop_flag = False
            self.t = threading.Thread(target=self._pump_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump_nvml(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, rs))
            except Exception:
                break
            time.sleep(0.002)

    def _stop_nvml(self):
        n = self.nvml
        self.stop_flag = True
        self.t.join(timeout=1)
        try:
            mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        except Exception:
            mx = None
        bits = {
            "hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        reasons = set()
        for _, rs in self.samples:
            for k, b in bits.items():
                if rs & b:
                    reasons.add(k)
        sm = [x for x, _ in self.samples]
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(mx) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
            "source": "nvml, 2 ms period, during the timed region",
        }

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if getattr(self, "nvml", None):
            return self._stop_nvml()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


WORKLOADS = {
    # name: (BASELINE.json config, default width, height, bounces)
    "terrain": ("configs[1]", 3840, 2160, 0),
    "sponza": ("configs[2]", 1920, 1080, 1),
    "large": ("configs[3]", 3840, 2160, 2),
    "edits": ("configs[4]", 3840, 2160, 0),
    "file": ("(a voxel-map cache file of the reference, --scene-file)", 1920, 1080, 1),
}
_SCENE_FILE = [None]
_WL = {"name": "terrain", "view": (6, 4), "cam": None, "label": None, "reference_ok": True}


def build_scene(workload="terrain", rank=0, world=1, barrier=None, large=(128, 24, 128, 512)):
    """-> (scene, sync records, stats); fills _WL (view extent, camera, label) for the chosen workload.
    rank / world / barrier: with several ranks on a node the big scene is generated once (rank 0, into /dev/shm) and mapped by the others."""
    from scenes import camera, terrain

    _WL["name"] = workload
    if workload in ("terrain", "edits"):
        scene = terrain.bench_terrain()
        _WL.update(view=(6, 4), cam=camera.Camera(), reference_ok=True,
                   label=f"{scene.get('name', 'terrain')}, reference camera (512,128,512) yaw 1.52 pitch -0.5")
    elif workload == "sponza":
        from scenes import models

        scene = models.sponza(2048)  # Main.cpp:42-47: Sponza.gltf voxelised into 2048^3 at the origin
        # 2048 x 1024 x 2048 view (the model is 860 voxels tall; the reference CPU view of 512 would cut it off)
        _WL.update(view=(6, 5), cam=camera.Camera(pos=(420.3, 160.2, 1010.7), yaw=1.5, pitch=-0.15), reference_ok=False,
                   label="bundled assets/models/Sponza voxelised into 2048^3, camera (420,160,1011) yaw 1.5 pitch -0.15")
    elif workload == "large":
        # BASELINE configs[3]: 4096-wide terrain with >= 10 GB of bricks: the reference's noise tree over 128 x ny x 128 sectors, raised by
        # `shift` voxels (everything under the surface is solid rock whose voxel ids come from the noise, as TerrainGenerator.cpp:21-23 has it)
        nx, ny, nz, shift = large
        if terrain.fastnoise_available():
            shared = "/dev/shm" if world > 1 and os.path.isdir("/dev/shm") else None
            try:
                scene = terrain.terrain_fastnoise_big(nx, ny, nz, y_shift=shift, shared_dir=shared, is_writer=rank == 0, wait=barrier)
            except OSError as e:  # /dev/shm too small: every rank generates its own copy
                print(f"[bench] shared scene failed ({e}); generating per rank", file=sys.stderr)
                scene = terrain.terrain_fastnoise_big(nx, ny, nz, y_shift=shift)
        else:
            scene = terrain.terrain_hash(nx, min(ny, 7), nz, seed=12345, emissive=False)
            shift = 0
        sy_log2 = 4 if ny <= 16 else 5
        _WL.update(view=(7, sy_log2), cam=camera.Camera(pos=(2048.0, 128.0 + shift, 2048.0)), reference_ok=False,
                   label=f"{scene.get('name', 'terrain')} in a 4096x{32 << sy_log2}x4096 view, camera (2048,{128 + shift},2048) yaw 1.52 pitch -0.5")
    elif workload == "file":
        # any "cvox 0004" file the reference wrote (e.g. its logs/voxels_2k_sponza.dat, Main.cpp:38-49), read by scenes/cvox.py
        from scenes import cvox

        if not _SCENE_FILE[0]:
            raise SystemExit("--workload file needs --scene-file <cvox file>")
        scene = cvox.load_cvox(_SCENE_FILE[0])
        keys = np.array(list(scene["sectors"].keys()) or [(0, 0, 0)])
        lo, hi = keys.min(0), keys.max(0)
        scene["sectors"] = {k: v for k, v in scene["sectors"].items() if min(k) >= 0 and k[0] < 128 and k[2] < 128 and k[1] < 64}
        cx, cy, cz = (float((lo[a] + hi[a] + 1) * 16) for a in range(3))
        _WL.update(view=(7, 6), cam=camera.Camera(pos=(cx + 0.3, cy + 0.2, cz + 0.7), yaw=1.5, pitch=-0.15), reference_ok=False,
                   label=f"{scene['name']} in the reference GPU renderer's 4096x2048x4096 view, camera at the centre of the populated box")
    else:
        raise SystemExit(f"unknown workload {workload}")
    return scene, terrain.scene_records(scene), terrain.scene_stats(scene)


def bench_frame(width, height, bounces, part_index=0, part_count=1, frame_no=1, flags=0):
    from scenes import camera
    from voxelrt_b200 import capi

    cam = _WL["cam"] or camera.Camera()  # Main.cpp:76-78
    proj, inv, wo, frac = cam.matrices(width, height)
    flags |= _WL.get("frame_flags", 0)  # --glsl: every frame of the run is a VRT_FRAME_GLSL frame
    return capi.make_frame(width, height, inv, proj, wo, frac, frame_no=frame_no, bounces=bounces, flags=flags, part_index=part_index, part_count=part_count)


def _bind_to_gpu_numa_node(index):
    """Pins this process to the CPUs of the NUMA node GPU `index` hangs off (sysfs); returns the node or None when the platform hides it."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def host_threads():
    """Hardware threads this process may use.  NOT OMP_NUM_THREADS: torch.distributed.run exports OMP_NUM_THREADS=1 to its
    workers, which in round 1 silently turned the N > 1 reference arm into a one-thread run."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_arm(args, scene, recs, want_ref=True, shipped_flags=True):
    """The CPU path on ALL host threads (thread count passed explicitly).  -> (run() -> seconds, kind, cores, rays, build note)
    kind "reference" = the reference's own CpuRenderer.cpp (oracle/_ref), timed in the build with the flags the reference ships
    with (-O3 -march=native -ffast-math, src/CMakeLists.txt:36) when that library is present; "port" = the oracle restatement."""
    from oracle import pyoracle

    kind, runner, build = "port", None, "oracle/vrt_oracle.c: gcc -O2 -ffp-contract=off -fno-fast-math, OpenMP"
    cores = host_threads()
    if want_ref and _WL["reference_ok"]:
        try:
            from oracle import refharness

            fast = shipped_flags and refharness.available(True)
            if refharness.available(fast):
                runner = refharness.RefRenderer(scene, recs, shipped_flags=fast, threads=cores)
                kind = "reference"
                build = ("reference CpuRenderer.cpp, g++ -O3 -march=native -ffast-math (the flags it ships with), OpenMP over tile rows" if fast else
                         "reference CpuRenderer.cpp, g++ -O2 -march=native -fno-fast-math (parity build), OpenMP over tile rows")
        except Exception as e:  # pragma: no cover
            print(f"[bench] oracle/_ref unavailable ({e}); timing the oracle port", file=sys.stderr)
    w, h = args.width, args.height
    rays = w * h * (1 + args.bounces)
    if runner is None:
        orc = pyoracle.OracleMap(*_WL["view"])
        orc.set_palette(scene["palette"])
        orc.sync(recs)
        if args.bounces:
            from scenes import shading

            orc.set_blue_noise(shading.load_blue_noise()[0])
            d, t, _ = shading.load_sky()
            orc.set_sky(d, t)
        frame = bench_frame(w, h, args.bounces)

        def run():
            t0 = time.perf_counter()
            orc.render(frame, want_aux=False, threads=cores)
            return time.perf_counter() - t0

    else:

        def run():
            return runner.render_seconds(w, h, args.bounces)

    return run, kind, cores, rays, build


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene, recs, sstats = build_scene(args.workload)
    run, kind, cores, rays, build = cpu_arm(args, scene, recs)
    for _ in range(args.warmup):
        run()
    times = [run() for _ in range(args.steps)]
    ms = 1000.0 * sum(times) / len(times)
    val = rays / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference",
        "metric": "Mrays/s primary+secondary, brickmap traversal",
        "value": val,
        "unit": "Mrays/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, scene, sstats),
        "cpu_baseline": {
            "value": val,
            "unit": "Mrays/s",
            "cores": cores,
            "kind": kind,
            "build": build,
            "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
            "sample": f"{args.steps} full {args.width}x{args.height} frames, {rays} rays each, on {cores} host threads (explicit count; the launcher's OMP_NUM_THREADS is ignored)",
        },
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, scene, sstats):
    return {
        "workload": f"BASELINE {WORKLOADS[args.workload][0]}: {_WL['label']}, "
        f"{args.width}x{args.height} primary rays + normal/material/depth G-buffer, bounces={args.bounces}"
        + ((f", {args.edits} random voxel edits" if args.edit_mode == "random" else ", one brush stroke (capsule r=30, fill/erase/replace)") + " + vrt_sync before every frame" if args.workload == "edits" else ""),
        "width": args.width,
        "height": args.height,
        "bounces": args.bounces,
        "renderer": ("the reference's GPU renderer: Shaders/VoxelRender.comp per pixel (VRT_FRAME_GLSL" + (", anisotropic LODs" if args.glsl_aniso else "") +
                     "), rays counted like GpuRenderer.cpp:292-296 = pixels * (bounces + 1) * 2; parity with the GLSL unpinned") if getattr(args, "glsl", False)
        else "the reference's CPU renderer: RenderRow per 4x4 packet (CpuRenderer.cpp:326-402)",
        "bricks": sstats["bricks"],
        "sectors": sstats["sectors"],
        "l2": "flushed between timed frames (256 MiB memset outside the event-timed region)" if args.gpus == 1 else
              "N>1: frames pipelined back to back, brickmap L2-resident (no flush possible inside the pipelined region; see gather.value_frame_latency_flushed_l2_owner0)",
        "parallelism": f"screen tiles 32x32 round-robin over {args.gpus} GPU(s), brickmap replicated" if args.gpus == 1 else
                       f"8-pixel screen bands round-robin over {args.gpus} GPUs, brickmap replicated, bands gathered over NVLink",
    }


def roofline_block(args, world, alg_bytes, ms_per_step, peak, peak_src, m, my_primary, launches):
    """The path is NOT bandwidth-bound on the configs whose masks fit the 126 MB L2 (all but configs[3]'s bounce rays): ncu shows the frame
    kernels limited by instruction ISSUE (70-80 % of the issue slots busy, ALU pipe ~60 %), DRAM traffic a few % of the algorithmic bytes.
    `achieved / peak / frac` stay what the contract defines — algorithmic bytes (8 I_s + 8 I_c + 9 H + 16 P, with the REFERENCE's iteration
    counts, which the metrics build of the kernel counts and tests tie to the oracle) over the measured HBM copy peak, a work-equivalent
    figure — and `bound` says what really limits the kernel, with the ncu numbers of the committed capture beside it."""
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    ncu = _ncu_summary(args)
    kernel = ("vrt::k_render<false,true,%s>" % ("true" if world > 1 else "false")) if args.bounces == 0 else "vrt::k_wave_primary + k_wave_trace + k_wave_shade (or k_render<false,false,...>: self-tuned)"
    if getattr(args, "glsl", False):
        kernel = "vrt::k_render_glsl (no traversal counters for the GLSL walk: the algorithmic bytes below are the 16 B/px G-buffer only)"
        ncu = None
    return {
        "bound": "issue",
        "bound_note": "instruction-issue bound (see `ncu`); achieved/peak/frac = algorithmic bytes over the measured HBM copy peak, a work-equivalent figure",
        "kernel": kernel,
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": int(alg_bytes),
        "bytes_per_ray": alg_bytes / max(1, my_primary * (1 + args.bounces)),
        "traffic": None if ncu is None else ncu.get("dram_bytes"),
        "ncu": ncu,
        "launches_per_step": launches,
        "counters": {"rays": m.rays, "reference_iterations": m.iters, "sector_fetches": m.sector_fetches, "cell_fetches": m.cell_fetches, "hits": m.hits, "capped": m.capped},
    }


def present_block(args, ctx, frame, stream, fb, w, h, peak):
    """--present: the step after the path — the reference's GBuffer (tile blit + reprojection + SVGF + tone-mapped present,
    include/voxelrt_b200_post.h) on the frames the bench just traced.  Device-timed with CUDA events on the launching
    stream (camera panning a little every frame so the reprojection does real work), plus the trace-to-window time
    through vrt_gbuffer_render_present with a host RGBA8 buffer.  Algorithmic bytes: DESIGN.md §10."""
    import torch

    from scenes import camera as _camera
    from voxelrt_b200 import capi, post

    gb = post.GBuffer(torch.cuda.current_device())
    gb.set_passes(args.present_passes)
    d_rgba = torch.zeros(w * h, dtype=torch.int32, device="cuda")
    cams = []
    base = _WL["cam"] or _camera.Camera()  # Main.cpp:76-78
    for f in range(args.warmup + args.steps):
        cam = _camera.Camera(pos=(base.pos[0] + 0.25 * f, base.pos[1], base.pos[2] + 0.125 * f), yaw=base.yaw, pitch=base.pitch)
        proj, inv, wo, frac = cam.matrices(w, h)
        cams.append((capi.make_frame(w, h, inv, proj, wo, frac, frame_no=f + 1, bounces=args.bounces), post.make_camera(w, h, proj, inv, cam.pos)))
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for f, (fr, gc) in enumerate(cams):
        ctx.render_device(fr, fb.data_ptr(), None, stream.cuda_stream)
        gb.set_camera(gc)
        k = f - args.warmup
        if k >= 0:
            evs[k][0].record(stream)
        gb.denoise_present_device(fb.data_ptr(), d_rgba.data_ptr(), stream.cuda_stream)
        if k >= 0:
            evs[k][1].record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    launches = gb.last_launches()
    npx = w * h
    n = args.present_passes
    alg = npx * (58 + (49 + 32 * n if n else 0) + 28)
    # trace -> window through the one-call entry point, host RGBA8 buffer (4 B/px D2H instead of 16)
    gb.set_camera(cams[0][1])
    gb.render_present(ctx, cams[0][0])  # untimed: allocates the GBuffer's own tile / image buffers
    t0 = time.perf_counter()
    for fr, gc in cams[: args.steps]:
        gb.set_camera(gc)
        gb.render_present(ctx, fr)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    gb.close()
    cpu = None
    if not args.no_cpu:
        # reported baseline, not a target: the image-space ORACLE (a scalar C port of the shaders, OpenMP over rows) on one traced frame —
        # the reference itself runs this step as GLSL on a GPU and has no CPU implementation of it
        from oracle import pypostoracle

        tiles_host = fb.cpu().numpy()
        po = pypostoracle.PostOracle(w, h)
        po.set_passes(n)
        proj, inv, _, _ = base.matrices(w, h)
        po.set_camera(proj, inv, base.pos)
        t0 = time.perf_counter()
        po.denoise_present(tiles_host)
        cpu = {"ms_per_frame": (time.perf_counter() - t0) * 1e3, "kind": "port", "cores": os.cpu_count(),
               "sample": f"one {w}x{h} frame without history (every pixel takes the 7x7 variance pass), {n} passes"}
    return {
        "cpu_baseline": cpu,
        "what": "CopyTiledFramebuffer + Reproject + Filter(-1, 0..%d) + GBufferBlit as CUDA kernels on the traced frame" % (n - 1),
        "passes": n,
        "ms_per_frame": ms,
        "launches_per_frame": launches,
        "algorithmic_bytes_per_frame": alg,
        "achieved_GB_per_s": alg / (ms * 1e-3) / 1e9,
        "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak,
        "trace_to_window_ms": e2e_ms,
        "trace_to_window_d2h_bytes": npx * 4,
        "timing": "CUDA events on the launching stream around the %d launches of each frame, warm L2" % launches,
    }


def run_b200(args):
    import torch
    import torch.distributed as dist

    from voxelrt_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    def _barrier():
        if world > 1:
            dist.barrier()

    scene, recs, sstats = build_scene(args.workload, rank, world, _barrier, large=(128, args.large_y, 128, args.large_shift))
    cap = 1 << 18
    while cap < sstats["bricks"] + 4096:
        cap <<= 1
    ctx = capi.Context(*_WL["view"], device=local, initial_brick_capacity=cap)
    ctx.set_palette(scene["palette"])
    # whole-scene upload through vrt_sync (pinned staging, one H2D copy, upload / header / box kernels), timed once
    rec_arr, rec_keep, rec_n = capi.make_records(recs)
    torch.cuda.synchronize()
    t_up = time.perf_counter()
    ctx.sync_records(rec_arr, rec_n)
    torch.cuda.synchronize()
    t_up = time.perf_counter() - t_up
    up_stats = ctx.stats()
    # the same records once more (every brick dirty again, nothing allocated or moved): the steady-state rate of the upload path
    t_re = None
    if sstats["bricks"] <= 4_000_000:
        t_re = time.perf_counter()
        ctx.sync_records(rec_arr, rec_n)
        torch.cuda.synchronize()
        t_re = time.perf_counter() - t_re
    residency = {"scene_upload_ms": t_up * 1e3, "scene_upload_note": "first sync: includes the one-time pinned / device staging allocation",
                 "scene_upload_GB_per_s": up_stats.bytes_uploaded / t_up / 1e9,
                 "bricks": int(up_stats.bricks_uploaded), "h2d_bytes": int(up_stats.bytes_uploaded),
                 "reupload_ms": None if t_re is None else t_re * 1e3, "reupload_GB_per_s": None if t_re is None else up_stats.bytes_uploaded / t_re / 1e9,
                 "device_bytes": int(ctx.stats().device_bytes)}
    del rec_keep
    if args.bounces:
        from scenes import shading

        ctx.set_blue_noise(shading.load_blue_noise()[0])
        d, t, _ = shading.load_sky()
        ctx.set_sky(d, t)
    if args.wavefront is not None:
        ctx.set_option("wavefront", args.wavefront)

    w, h = args.width, args.height
    npx = w * h
    rays_frame = npx * (1 + args.bounces)
    if args.glsl:
        # the GPU renderer's frame shader (VRT_FRAME_GLSL) and its own ray accounting: (bounces + 1) * 2 rays per pixel, "sun" included
        # (GpuRenderer.cpp:292-296) — the shader casts 1 + [b > 0] + b + min(b, 2) at most
        _WL["frame_flags"] = capi.VRT_FRAME_GLSL | (capi.VRT_FRAME_GLSL_ANISOTROPIC if args.glsl_aniso else 0)
        rays_frame = npx * (1 + args.bounces) * 2
    # primary-only frames: 8 B/px (VRT_FRAME_COMPACT: albedo + normal, depth; the irradiance of such a frame is the constant 1.0 and is
    # not moved) over NVLink and PCIe; the device-resident N = 1 frame keeps the reference's full 16 B/px tiles
    compact = args.bounces == 0 and args.compact and not args.glsl
    xfer_px = 8 if compact else 16
    xflag = capi.VRT_FRAME_COMPACT if compact else 0
    fb = torch.zeros(npx * 4, dtype=torch.int32, device="cuda")  # 16 B/px
    out_ptr = fb.data_ptr()
    # N > 1: pipelined tile gather over NVLink (vrt_render_gather).  Every rank renders its 8-pixel bands into its own
    # buffer; one strided copy per frame then moves them into the presenting GPU's framebuffer while the next frame is
    # traced.  The presenting GPU rotates with the frame number (frame f is assembled on GPU f % N — one encoder /
    # display head per GPU), so no single NVLink port has to swallow 7/8 of every frame.
    owner_ptrs, local_fbs = None, None
    if world > 1:
        handle, own_ptr = ctx.fb_export(npx * xfer_px)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        owner_ptrs = [own_ptr if r == rank else ctx.fb_import(handles[r]) for r in range(world)]
        local_fbs = [torch.zeros(npx * xfer_px // 4, dtype=torch.int32, device="cuda") for _ in range(capi.VRT_GATHER_DEPTH)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # a real (non-legacy) stream: handle 0 would mean "the context's own stream" to the C ABI and
    # torch events recorded on the legacy stream would not bracket the kernel
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    # N > 1: consecutive frames rotate over a few streams, so the tail of frame f (its last, longest warp tiles: a capped
    # ray keeps its warp busy for ~40 us, most of a 50 us frame at 8 GPUs) overlaps the next frames on the SMs
    n_streams = int(os.environ.get("VRT_BENCH_STREAMS", "8"))
    streams = [stream] + [torch.cuda.Stream() for _ in range(n_streams - 1)] if world > 1 else [stream]
    join_evs = [torch.cuda.Event() for _ in streams]

    def join_streams():  # `stream` waits for everything issued on the other streams and for the last gather
        for s_, e_ in zip(streams[1:], join_evs[1:]):
            e_.record(s_)
            stream.wait_event(e_)
        ctx.gather_wait(stream.cuda_stream)

    def fork_streams():  # the other streams start after everything issued on `stream` so far
        join_evs[0].record(stream)
        for s_ in streams[1:]:
            s_.wait_event(join_evs[0])
    frame = bench_frame(w, h, args.bounces, part_index=rank, part_count=world, flags=(capi.VRT_FRAME_PART_ROWS | xflag) if world > 1 else 0)
    frame_i = [0]

    # workload "edits" (BASELINE configs[4]): every frame is preceded by a batch of voxel edits whose dirty bricks are
    # uploaded by vrt_sync (pinned staging + one H2D copy + upload / header / box kernels) inside the timed region; the
    # host-side world editing (the reference's VoxelMap::Set) is precomputed, one record batch per frame
    edit_batches, edit_i, edit_stats = None, [0], {"bricks": 0, "bytes": 0, "syncs": 0, "launches": 0}
    if args.workload == "edits":
        from scenes import edits as _edits

        n_frames = 3 * args.steps + args.warmup + 8
        if args.edit_mode == "brush":  # the reference's brush: capsule strokes of radius 30 (Brush.cpp), fill / erase / replace
            frames_recs, _ = _edits.brush_stroke_frames(scene, n_frames, seed=1)
        else:
            frames_recs, _ = _edits.random_edit_frames(scene, n_frames, args.edits, seed=1)
        edit_batches = [capi.make_records(r) for r in frames_recs]

    def step():
        if edit_batches is not None:
            arr, keep, n = edit_batches[edit_i[0] % len(edit_batches)]
            edit_i[0] += 1
            ctx.sync_records(arr, n)
            st_ = ctx.stats()
            edit_stats["bricks"] += st_.bricks_uploaded
            edit_stats["bytes"] += st_.bytes_uploaded
            edit_stats["syncs"] += 1
            edit_stats["launches"] += st_.last_launches
        if world > 1:
            i = frame_i[0]
            frame_i[0] += 1
            owner = {"rotate": i % world, "self": rank, "zero": 0}[os.environ.get("VRT_BENCH_OWNER", "rotate")]
            # the presenting rank renders its own bands straight into its framebuffer (no local copy)
            local_ptr = owner_ptrs[owner] if owner == rank else local_fbs[i % capi.VRT_GATHER_DEPTH].data_ptr()
            ctx.render_gather(frame, local_ptr, owner_ptrs[owner], streams[i % len(streams)].cuda_stream)
        else:
            ctx.render_device(frame, out_ptr, None, stream.cuda_stream)

    # traversal counters of this frame (untimed, metrics build of the same kernel)
    ctx.set_option("metrics", 1)
    step()
    torch.cuda.synchronize()
    m = ctx.metrics()
    ctx.set_option("metrics", 0)
    from voxelrt_b200 import partition

    step()
    torch.cuda.synchronize()
    # frames with bounces: the context times its two forms (one thread per pixel / wavefront passes) on the first four frames of a new
    # (size, bounces, scene) combination and keeps the faster one; let that settle before anything is measured
    for _ in range(8):
        if not (int(ctx.stats().bounce_form) & 0x100):
            break
        step()
        torch.cuda.synchronize()
    step()
    torch.cuda.synchronize()
    bounce_form = {0: None, 1: "one thread per pixel", 2: "wavefront passes"}[int(ctx.stats().bounce_form) & 0xFF]
    if world > 1:
        forms = [None] * world
        dist.all_gather_object(forms, bounce_form)
        bounce_form = forms[0] if len(set(forms)) == 1 else forms
    ctx_launches = int(ctx.stats().last_launches)  # kernels one step launches (1 for a primary frame; 1 + 3 per bounce level for a wavefront frame)
    my_primary = partition.pixels_of_rank(w, h, rank, world, rows=world > 1)
    alg_bytes = 8 * m.sector_fetches + 8 * m.cell_fetches + 9 * m.hits + 16 * my_primary

    for _ in range(args.warmup):
        if edit_batches is None:
            flush.zero_()
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for k in edit_stats:
        edit_stats[k] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    if world == 1:
        for a, b in evs:
            if edit_batches is None:
                flush.zero_()
            a.record(stream)
            step()
            b.record(stream)
    else:
        # N > 1: the K frames go back to back (trace of frame f+1 overlaps the gather of frame f); one event pair brackets
        # all of them INCLUDING the last gather.  No L2 flush is possible inside a pipelined region: the brickmap working
        # set (9 MB) is L2-resident, which the N=1 line reports separately as value_warm_l2 (+0.5 %).
        a, b = evs[0]
        a.record(stream)
        fork_streams()
        for _ in range(args.steps):
            step()
        join_streams()
        b.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    total_ms = sum(a.elapsed_time(b) for a, b in evs) if world == 1 else evs[0][0].elapsed_time(evs[0][1])
    render_only_ms = total_ms / args.steps
    timed_edit_stats = dict(edit_stats)
    if edit_batches is not None:
        # edit -> visible: host-side vrt_sync (staging, H2D, upload kernels) + the frame; the events above see the frame only
        total_ms = t_wall * 1000.0
    # warm-L2 variant (steady-state renderer: brickmap stays in the 126 MB L2), informational
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    if world > 1:
        fork_streams()
    for _ in range(args.steps):
        step()
    if world > 1:
        join_streams()
    b.record(stream)
    torch.cuda.synchronize()
    warm_ms = a.elapsed_time(b) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    # the step-by-step loop (no empty-box macro steps), informational
    ctx.set_option("macro_steps", 0)
    step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    if world > 1:
        fork_streams()
    for _ in range(args.steps):
        step()
    if world > 1:
        join_streams()
    b.record(stream)
    torch.cuda.synchronize()
    stepwise_ms = a.elapsed_time(b) / args.steps
    ctx.set_option("macro_steps", 1)

    trace_only_ms, latency_ms, gather_ok, replicas_ok = None, None, None, None
    if world > 1:
        def local_step(k=0):  # the same split without the NVLink gather: every rank keeps its bands
            ctx.render_device(frame, local_fbs[k % len(local_fbs)].data_ptr(), None, streams[k % len(streams)].cuda_stream)

        local_step()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fork_streams()
        for k in range(args.steps):
            local_step(k)
        join_streams()
        b.record(stream)
        torch.cuda.synchronize()
        # frame LATENCY: one frame at a time, L2 flushed, all bands gathered on rank 0 before the next frame starts
        lat = 0.0
        for _ in range(args.steps):
            dist.barrier()
            flush.zero_()
            c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c.record(stream)
            ctx.render_gather(frame, local_fbs[0].data_ptr(), owner_ptrs[0], stream.cuda_stream)
            ctx.gather_wait(stream.cuda_stream)
            d.record(stream)
            torch.cuda.synchronize()
            lat += c.elapsed_time(d)
        lt = torch.tensor([a.elapsed_time(b) / args.steps, lat / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.MAX)
        trace_only_ms, latency_ms = float(lt[0]), float(lt[1])
        # the frame assembled in rank 0's memory must equal the frame rank 0 renders alone
        dist.barrier()
        if rank == 0:
            # (edits workload: every rank has applied the same edit batches, so rank 0 alone must still render the gathered frame)
            solo = torch.zeros(npx * xfer_px // 4, dtype=torch.int32, device="cuda")
            ctx.render_device(bench_frame(w, h, args.bounces, flags=xflag), solo.data_ptr(), None, stream.cuda_stream)
            torch.cuda.synchronize()

            class _DevPtr:  # view rank 0's exported framebuffer as a tensor
                __cuda_array_interface__ = {"shape": (npx * xfer_px // 4,), "typestr": "<i4", "data": (int(owner_ptrs[0]), False), "version": 2}

            gathered = torch.as_tensor(_DevPtr(), device="cuda")
            gather_ok = bool(torch.equal(solo, gathered))
        dist.barrier()
        # replicas: the resident brickmap of every rank must be the same (same records, same allocator decisions) — compare a digest of
        # the device state of a sample of sectors plus the allocator's statistics across the ranks
        import hashlib as _hl

        hsh = _hl.sha256()
        keys = sorted(scene["sectors"].keys())
        for key in keys[:: max(1, len(keys) // 64)][:64]:
            m_, base_, bricks_, cells_ = ctx.read_sector(*key)
            hsh.update(int(m_).to_bytes(8, "little") + int(base_).to_bytes(4, "little") + bricks_.tobytes() + cells_.tobytes())
        st_ = ctx.stats()
        hsh.update(f"{st_.resident_bricks},{st_.resident_sectors},{st_.free_ranges}".encode())
        digests = [None] * world
        dist.all_gather_object(digests, hsh.hexdigest())
        replicas_ok = len(set(digests)) == 1
    tt = torch.tensor([total_ms, warm_ms], dtype=torch.float64, device="cuda")
    ab = torch.tensor([float(alg_bytes)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, warm_ms = float(tt[0]), float(tt[1])
    ms_per_step = total_ms / args.steps
    value = rays_frame / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the host-buffer ABI call (vrt_render), wall clock.  Frame constants in, G-buffer out into page-locked HOST memory.
    # N = 1: the whole frame into a pinned buffer.  N > 1: ONE host frame shared by the ranks of the node (POSIX shared memory, page-locked
    # in every process): each rank's vrt_render traces its 8-pixel bands and copies them to their place in that frame over its own PCIe
    # link, so a complete frame exists on the host after every step and the copy bandwidth scales with the GPUs.
    shm = None
    if world == 1:
        host_out = torch.empty(npx * xfer_px // 4, dtype=torch.int32).pin_memory()
        host_ptr = host_out.data_ptr()
    else:
        from multiprocessing import shared_memory

        names = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=npx * xfer_px)
            names[0] = shm.name
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=names[0])
            try:  # (rank 0 owns and unlinks the segment; keep the other ranks' resource trackers from complaining about it at exit)
                from multiprocessing import resource_tracker

                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        host_np = np.ndarray((npx * xfer_px // 4,), dtype=np.int32, buffer=shm.buf)
        host_ptr = host_np.ctypes.data
        # NUMA placement (best effort): every rank runs on the CPUs next to its GPU and first-touches ITS bands of the frame before
        # anybody pins the pages, so that a band's pages live on the socket whose PCIe root the copy arrives at
        numa = _bind_to_gpu_numa_node(local)
        band_words = (w // 4) * 2 * (xfer_px * 16) // 4
        for b_ in range(rank, (h + 7) // 8, world):
            host_np[b_ * band_words : (b_ + 1) * band_words] = 0
        dist.barrier()
        rc = torch.cuda.cudart().cudaHostRegister(host_ptr, npx * xfer_px, 0)
        if int(rc) != 0:
            print(f"[bench] cudaHostRegister of the shared host frame failed ({rc}); copies go through pageable memory", file=sys.stderr)
        dist.barrier()
    e2e_frame = bench_frame(w, h, args.bounces, part_index=rank, part_count=world, flags=(capi.VRT_FRAME_PART_ROWS if world > 1 else 0) | xflag)

    def e2e_step(fr=e2e_frame, ptr=None):
        if edit_batches is not None:  # configs[4]: the frame's edits are part of the step
            arr, keep, n = edit_batches[edit_i[0] % len(edit_batches)]
            edit_i[0] += 1
            ctx.sync_records(arr, n)
        st = ctx.lib.vrt_render(ctx.h, C.byref(fr), host_ptr if ptr is None else ptr, None)
        if st != 0:
            ctx._chk(st)

    for _ in range(max(2, args.warmup // 2)):
        e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    if world > 1:
        dist.barrier()  # the host frame is complete when every rank has delivered its bands
    e2e_s = time.perf_counter() - t0
    et = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_ms = float(et[0]) * 1000.0 / args.steps
    e2e_val = rays_frame / (e2e_ms * 1e-3) / 1e6
    e2e_host_frame_ok = None
    if world > 1:
        if rank == 0 and edit_batches is None:  # the frame the ranks assembled on the host == the frame rank 0 delivers alone
            solo_host = torch.empty(npx * xfer_px // 4, dtype=torch.int32).pin_memory()
            e2e_step(bench_frame(w, h, args.bounces, flags=xflag), solo_host.data_ptr())
            n_diff = int(np.count_nonzero(solo_host.numpy() != host_np))
            e2e_host_frame_ok = n_diff == 0
            if n_diff:
                print(f"[bench] host frame differs from the single-GPU frame in {n_diff} words", file=sys.stderr)
        dist.barrier()
        torch.cuda.cudart().cudaHostUnregister(host_ptr)
        del host_np
        shm.close()
        if rank == 0:
            shm.unlink()
    # the reference's full 16 B/px tile framebuffer through the same call, for comparison (N = 1)
    e2e_full_val = None
    if world == 1 and compact:
        host_full = torch.empty(npx * 4, dtype=torch.int32).pin_memory()
        full_frame = bench_frame(w, h, args.bounces)
        for _ in range(2):
            e2e_step(full_frame, host_full.data_ptr())
        t0 = time.perf_counter()
        for _ in range(max(3, args.steps // 2)):
            e2e_step(full_frame, host_full.data_ptr())
        e2e_full_val = rays_frame / ((time.perf_counter() - t0) / max(3, args.steps // 2)) / 1e6

    if rank == 0:
        peak, peak_src = _peaks()
        kernel_s = ms_per_step * 1e-3
        achieved = alg_bytes / kernel_s / 1e9
        line = {
            "metric": "Mrays/s primary+secondary, brickmap traversal",
            "value": value,
            "unit": "Mrays/s",
            "n_gpus": n_gpus,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, scene, sstats),
            "value_warm_l2": rays_frame / (warm_ms * 1e-3) / 1e6,
            "value_stepwise_warm_l2": rays_frame / (stepwise_ms * 1e-3) / 1e6,
            "wall_s_timed_region": t_wall,
            "clocks": clocks,
            "e2e": {
                "value": e2e_val,
                "unit": "Mrays/s",
                "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": C.sizeof(capi.VrtFrame) * world + (int(timed_edit_stats["bytes"] / max(1, timed_edit_stats["syncs"])) * world if edit_batches is not None else 0),
                "d2h_bytes_per_step": npx * xfer_px,
                "payload": ("VRT_FRAME_COMPACT, 8 B/px: albedo + normal and depth; the irradiance words of a primary-only frame are the constant 1.0 "
                            "(CpuRenderer.cpp:379-381) and are not moved" if compact else "the reference's 16 B/px tile framebuffer"),
                "value_full_16B_per_px": e2e_full_val,
                "api": "vrt_render (host buffers, page-locked output)" + ("" if world == 1 else
                       "; ONE host frame in shared page-locked memory, every rank delivers its bands into it over its own PCIe link"),
                "host_frame_equal_to_single_gpu_frame": e2e_host_frame_ok,
                "numa_node_of_rank0": numa if world > 1 else None,
            },
            "gpu_launches": args.steps * ctx_launches + timed_edit_stats["launches"],
            "bounce_form": bounce_form,
            "residency": residency,
            "edits": None if edit_batches is None else {
                "mode": args.edit_mode,
                "voxel_edits_per_frame": args.edits if args.edit_mode == "random" else None,
                "dirty_bricks_per_frame": timed_edit_stats["bricks"] / max(1, timed_edit_stats["syncs"]),
                "h2d_bytes_per_frame": timed_edit_stats["bytes"] / max(1, timed_edit_stats["syncs"]),
                "render_only_ms": render_only_ms,
                "edit_to_visible_ms": ms_per_step,
                "timing": "wall clock over the timed loop (vrt_sync is a blocking host call), no L2 flush",
            },
            "gather": None if world == 1 else {
                "how": "vrt_render_gather: 8-pixel bands rendered locally, one strided D2D copy per frame into the presenting GPU's framebuffer "
                       "(CUDA IPC over NVLink; copy engines, not the SMs) on a copy stream, overlapped with the next frames; presenting GPU = frame % N; "
                       "NCCL carries only the control plane (handles, barriers, timing reductions)",
                "verified_equal_to_single_gpu_frame": gather_ok,
                "replicas_identical": replicas_ok,
                "payload_bytes_per_pixel": xfer_px,
                "value_trace_only_warm_l2": rays_frame / (trace_only_ms * 1e-3) / 1e6,
                "value_frame_latency_flushed_l2_owner0": rays_frame / (latency_ms * 1e-3) / 1e6,
                "frame_latency_ms": latency_ms,
                "bytes_over_nvlink_per_frame": npx * xfer_px * (world - 1) // world,
            },
            "roofline": roofline_block(args, world, alg_bytes, ms_per_step, peak, peak_src, m, my_primary, ctx_launches),
        }
        if args.present and n_gpus == 1:
            # the step after the path; reported next to the headline, never allowed to take the bench line down with it
            try:
                line["present"] = present_block(args, ctx, frame, stream, fb, w, h, peak)
            except Exception as e:  # noqa: BLE001
                line["present"] = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_cpu and n_gpus == 1:
            run, kind, cores, rays, build = cpu_arm(args, scene, recs)
            run()
            times, t_begin = [], time.perf_counter()
            while len(times) < 3 or (time.perf_counter() - t_begin < 10.0 and len(times) < 20):
                times.append(run())
            best = min(times)
            line["cpu_baseline"] = {
                "value": rays / best / 1e6,
                "unit": "Mrays/s",
                "cores": cores,
                "kind": kind,
                "build": build,
                "sample": f"best of {len(times)} full {w}x{h} frames ({rays} rays each) on {cores} host threads",
            }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="terrain", choices=sorted(WORKLOADS), help="terrain = BASELINE configs[1] (the headline); the others are the remaining configs")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--bounces", type=int, default=None)
    ap.add_argument("--edits", type=int, default=4096, help="workload 'edits': voxel edits per frame")
    ap.add_argument("--edit-mode", default="random", choices=["random", "brush"], help="workload 'edits': uniform single-voxel edits, or the reference's brush strokes")
    ap.add_argument("--present", dest="present", action="store_true", default=True, help="also time the step after the path (the reference's GBuffer: denoise + present) and add a 'present' block (default at N = 1)")
    ap.add_argument("--no-present", dest="present", action="store_false", help="skip the 'present' block")
    ap.add_argument("--present-passes", type=int, default=5, help="GBuffer::NumDenoiserPasses for --present (0..5)")
    ap.add_argument("--scene-file", default=None, help="workload 'file': a cvox 0004 voxel-map file written by the reference (or by scenes/cvox.py)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--wavefront", type=int, default=None, choices=[0, 1, 2], help="frames with bounces: 0 one thread per pixel, 1 wavefront passes, 2 self-tuning (library default)")
    ap.add_argument("--glsl", action="store_true", help="render VRT_FRAME_GLSL frames: the reference's GPU renderer (Shaders/VoxelRender.comp per pixel: sun shadow ray, coarse bounce rays); rays counted like GpuRenderer.cpp:292-296; the traversal counters of the roofline block do not exist for this walk")
    ap.add_argument("--glsl-aniso", action="store_true", help="with --glsl: u_UseAnisotropicLods")
    ap.add_argument("--no-compact", dest="compact", action="store_false", default=True, help="move the full 16 B/px tiles of primary-only frames over NVLink / PCIe instead of the 8 B/px that carry information")
    ap.add_argument("--large-y", type=int, default=24, help="workload 'large': sectors along y (24 with --large-shift 512 = 10 GB of bricks)")
    ap.add_argument("--large-shift", type=int, default=512, help="workload 'large': voxels the terrain is raised by (solid rock below the surface)")
    args = ap.parse_args()
    _SCENE_FILE[0] = args.scene_file
    _, dw, dh, db = WORKLOADS[args.workload]
    args.width = dw if args.width is None else args.width
    args.height = dh if args.height is None else args.height
    args.bounces = db if args.bounces is None else args.bounces
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
