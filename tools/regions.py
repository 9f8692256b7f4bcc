#!/usr/bin/env python
"""Where the frame kernel's instructions go, by SOURCE region.

    python tools/regions.py [gpurun_out/prof_render_source.csv] [out.txt]

Joins the per-SASS-instruction counters of an `ncu --set full --import-source on` capture (exported with
`ncu -i rep --page source --csv`) with the line table of the very library that was profiled (`nvdisasm -g` on the
cubin inside voxelrt_b200/lib/libvoxelrt_b200.so, compiled -lineinfo) and sums executed warp instructions, thread
utilisation and stall samples per function of voxelrt_b200/csrc — with the hot loop split at its labels.
"""
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
import os

# (a capture must be joined with the line table of the very build that was profiled: VRT_REGIONS_LIB / VRT_REGIONS_SRC point at it when
# the tree has moved on since)
LIB = Path(os.environ.get("VRT_REGIONS_LIB", ROOT / "voxelrt_b200" / "lib" / "libvoxelrt_b200.so"))
SRC = Path(os.environ.get("VRT_REGIONS_SRC", ROOT / "voxelrt_b200" / "csrc"))


def line_table(kernel_substr):
    """-> {sass offset: (file name, line)} for the first kernel whose mangled name contains kernel_substr."""
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(LIB)], cwd=td, check=True, capture_output=True)
        table = {}
        for cubin in Path(td).glob("*.cubin"):
            txt = subprocess.run(["nvdisasm", "-gi", str(cubin)], capture_output=True, text=True).stdout
            on, cur, fresh = False, [], True
            for ln in txt.splitlines():
                if ln.startswith("//---") and ".text." in ln:
                    on = kernel_substr in ln
                    continue
                if not on:
                    continue
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:  # a chain of frames, innermost first, precedes the instructions it covers
                    if fresh:
                        cur, fresh = [], False
                    cur.append((Path(m.group(1)).name, int(m.group(2))))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
                if m and cur:
                    table[int(m.group(1), 16)] = list(cur)
                    fresh = True
            if table:
                return table
    return {}


def source_regions():
    """-> {file name: sorted [(first line, region name)]}: one region per function, cast_loop_fast split at its labels."""
    out = {}
    fn = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__ __device__)[^;{]*?\b(\w+)\s*\(")
    strip = re.compile(r"__launch_bounds__\s*\((?:[^()]|\([^()]*\))*\)")
    for f in SRC.glob("*.cu*"):
        regs, lines = [], f.read_text().splitlines()
        pending_template = None
        for i, ln in enumerate(lines, 1):
            if ln.startswith("template"):
                pending_template = i
            ln = strip.sub("", ln)
            m = fn.match(ln) or (pending_template == i - 1 and fn.match("template <> " + ln) if pending_template else None)
            if m and not ln.rstrip().endswith(";"):
                regs.append((pending_template if pending_template == i - 1 else i, m.group(1)))
        if f.name == "vrt_device.cuh":
            marks = [("L_iter : {", "loop: position -> sector header -> brick bit"), ("if ((h.x | h.y) == 0u) {", "loop: empty / outside sector"),
                     ("if ((h.w & VRT_HDR_HASBOX)", "loop: empty-box macro jump attempt"), ("km = (((half << (sh & 0xAu)) & 0xCC00CC00u) == 0u) ? ~15 : ~7;", "loop: brick absent (lod 3/4)"),
                     ("} else {  // :146-158 brick present", "loop: brick present -> cell mask -> voxel bit / lod"), ("L_step:", "loop: step to the cell's far corner"),
                     ('asm volatile("");  // keeps the exits separate blocks', "loop exits")]
            for needle, name in marks:
                for i, ln in enumerate(lines, 1):
                    if needle in ln:
                        regs.append((i, name))
                        break
        out[f.name] = sorted(regs)
    return out


MAJOR = ("loop", "lean_trip", "k_wave_trace", "k_wave_shade", "k_wave_primary", "packet_lane", "packet_votes", "shade_bounce", "wave_", "bounce_direction",
         "cast_loop_generic", "cast_finish", "voxel_palette_id", "primary_ray", "shade_pixel", "store_pixel", "store_hit", "warp_tile_origin",
         "cast_ray", "cast_loop_fast", "render_warp_tile", "k_render", "sky_sample", "blue_noise", "sample_direction")


def _region_of_frame(regs, file, line):
    best = "(other)"
    for first, name in regs.get(file, []):
        if first <= line:
            best = name
        else:
            break
    return best if file in regs else f"({file})"


def region_of(regs, frames):
    """frames: inline chain, innermost first.  Small helpers (lop3_*, rcp_rn_normal, x86_min ...) are charged to the first
    enclosing major region, so the loop's numbers are not scattered over its helpers."""
    names = [_region_of_frame(regs, f, l) for f, l in frames]
    for n in names:
        if n.startswith(MAJOR):
            return n
    return names[0] if names else "(no line info)"


def report(kernel, hdr, data, src_name, regs):
    ia, ie, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    targs = kernel.split("(vrt::")[0]
    mname = re.search(r"vrt::(\w+)", kernel)
    base = mname.group(1) if mname else "k_render"
    # mangled template arguments: <(bool)0, (bool)1> -> ILb0ELb1EE, <(int)10> -> ILi10EE
    margs = "".join(f"L{'b' if t == 'bool' else 'i'}{v}E" for t, v in re.findall(r"\((bool|int)\)(\d+)", targs))
    sub = f"{len(base)}{base}I{margs}" if margs else f"{len(base)}{base}"
    table = line_table(sub)
    base_addr = int(data[0][ia], 16)
    agg, tot_e, tot_s = {}, 0, 0
    for r in data:
        off = int(r[ia], 16) - base_addr
        name = region_of(regs, table.get(off, []))
        a = agg.setdefault(name, [0, 0, 0, 0])
        e, t, s = int(r[ie]), int(r[it]), int(r[isamp])
        a[0] += e
        a[1] += t
        a[2] += s
        a[3] += 1
        tot_e += e
        tot_s += s
    lines = [f"kernel: {kernel}", f"capture: {src_name}; line table: nvdisasm -g of {LIB.name}", f"executed warp instructions: {tot_e}   stall samples: {tot_s}", "",
             f"{'source region':58s} {'SASS':>5s} {'warp inst':>12s} {'%inst':>6s} {'thr/inst':>8s} {'%samples':>8s}"]
    for name, (e, t, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if e == 0 and s == 0:
            continue
        lines.append(f"{name:58s} {n:5d} {e:12d} {100 * e / max(tot_e, 1):6.1f} {t / max(e, 1):8.1f} {100 * s / max(tot_s, 1):8.1f}")
    return "\n".join(lines) + "\n"


def main():
    src_csv = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "prof_render_source.csv"
    out = Path(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(open(src_csv)))
    # a capture of several launches repeats the ("Kernel Name", name) / header / instruction-rows block: one table per launch
    blocks, k = [], 0
    while k < len(rows):
        if rows[k] and rows[k][0] == "Kernel Name":
            kernel, hdr = rows[k][1], rows[k + 1]
            e = k + 2
            while e < len(rows) and not (rows[e] and rows[e][0] == "Kernel Name"):
                e += 1
            blocks.append((kernel, hdr, [r for r in rows[k + 2 : e] if r]))
            k = e
        else:  # single-launch export without the name row
            blocks.append(("k_render", rows[0], [r for r in rows[1:] if r]))
            break
    regs = source_regions()
    uniq, seen = [], set()
    for kernel, hdr, data in blocks:  # (ncu exports every launch twice: once per view of the source page)
        key = (kernel, data[0][0] if data else None, len(data), sum(int(r[hdr.index("Instructions Executed")]) for r in data))
        if data and key not in seen:
            seen.add(key)
            uniq.append((kernel, hdr, data))
    text = "\n".join(report(kernel, hdr, data, src_csv.name.lstrip("."), regs) for kernel, hdr, data in uniq)
    if out:
        out.write_text(text)
    print(text)


if __name__ == "__main__":
    main()
