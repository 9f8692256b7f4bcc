#!/usr/bin/env python
"""Turn a gpurun_out/ session (tools/gpu_round.sh) into the tracked summaries under profiles/.

    python tools/profiles.py [tag]          # tag defaults to r01
"""
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
METRICS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum l1tex__t_sector_hit_rate.pct
lts__t_sector_hit_rate.pct smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__grid_size launch__block_size
smsp__thread_inst_executed_per_inst_executed.ratio lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed lts__t_sectors.sum""".split()


def summarise(rep, name):
    """ncu report -> profiles/<name>_details.txt + <name>_raw_summary.csv (the metrics listed above, one column per launch)."""
    det = subprocess.run(["ncu", "-i", str(rep), "--page", "details"], capture_output=True, text=True).stdout
    (PROF / f"{name}_details.txt").write_text(det)
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, launches = rows[0], rows[1], rows[2:]
    extra = ["smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
    with open(PROF / f"{name}_raw_summary.csv", "w") as f:
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(launches))) + "\n")
        for m in METRICS + extra:
            i = next((j for j, h in enumerate(head) if h == m), None)
            if i is not None:
                f.write(f"{m},{units[i]}," + ",".join(r[i].replace(",", "") for r in launches) + "\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    if (OUT / "prof_bounce.ncu-rep").exists():
        summarise(OUT / "prof_bounce.ncu-rep", f"{tag}_ncu_k_render_bounce1")
        if (OUT / "prof_bounce_source.csv").exists():
            subprocess.run([sys.executable, str(ROOT / "tools" / "regions.py"), str(OUT / "prof_bounce_source.csv"), str(PROF / f"{tag}_ncu_k_render_bounce1_source_regions.txt")],
                           stdout=subprocess.DEVNULL)
    if (OUT / "prof_wave.ncu-rep").exists():
        summarise(OUT / "prof_wave.ncu-rep", f"{tag}_ncu_k_wave_trace_2bounces")
    if (OUT / "prof_atrous.ncu-rep").exists():  # the GBuffer step's filter kernel
        summarise(OUT / "prof_atrous.ncu-rep", f"{tag}_ncu_k_atrous")
    rep = OUT / "prof_render.ncu-rep"
    if rep.exists():
        det = subprocess.run(["ncu", "-i", str(rep), "--page", "details"], capture_output=True, text=True).stdout
        (PROF / f"{tag}_ncu_k_render_details.txt").write_text(det)
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        head, units, launches = rows[0], rows[1], rows[2:]
        col = {}
        with open(PROF / f"{tag}_ncu_k_render_raw_summary.csv", "w") as f:
            f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(launches))) + "\n")
            for m in METRICS:
                i = next((j for j, h in enumerate(head) if h == m), None)
                if i is None:
                    continue
                f.write(f"{m},{units[i]}," + ",".join(r[i].replace(",", "") for r in launches) + "\n")
                if m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[units[i]]
                    col[m + "#bytes"] = int(round(float(launches[0][i].replace(",", "")) * scale))
        b = json.loads((OUT / "bench.json").read_text().strip().splitlines()[-1]) if (OUT / "bench.json").exists() else None
        if b and "dram__bytes_read.sum#bytes" in col:
            key = f"k_render<false,true> {b['config']['width']}x{b['config']['height']} bounces={b['config']['bounces']}"
            t = json.loads((PROF / "traffic.json").read_text())
            t[key] = {"dram_read_bytes": col["dram__bytes_read.sum#bytes"], "dram_write_bytes": col["dram__bytes_write.sum#bytes"],
                      "capture": f"profiles/{tag}_ncu_k_render_raw_summary.csv launch0"}
            (PROF / "traffic.json").write_text(json.dumps(t, indent=2) + "\n")
    if (OUT / "prof_render_source.csv").exists():
        subprocess.run([sys.executable, str(ROOT / "tools" / "regions.py"), str(OUT / "prof_render_source.csv"), str(PROF / f"{tag}_ncu_k_render_source_regions.txt")],
                       stdout=subprocess.DEVNULL)
    for src, dst in (("bench.json", f"{tag}_bench_4k_primary.json"), ("bench_ref.json", f"{tag}_bench_reference_arm.json"),
                     ("launches.csv", f"{tag}_launches_4k_primary.csv"), ("bench_sponza.json", f"{tag}_bench_sponza_1080p_1bounce.json"),
                     ("bench_large.json", f"{tag}_bench_large_4k_2bounces.json"), ("bench_edits.json", f"{tag}_bench_edits_4k.json"), ("bench_edits_brush.json", f"{tag}_bench_edits_brush_4k.json"),
                     ("bench_2gpu.json", f"{tag}_bench_4k_primary_2gpu.json"), ("bench_4gpu.json", f"{tag}_bench_4k_primary_4gpu.json"),
                     ("bench_8gpu.json", f"{tag}_bench_4k_primary_8gpu.json"), ("macro_stats.log", f"{tag}_macro_stats.txt"),
                     ("launches_post.csv", f"{tag}_launches_post_4k.csv"), ("post_check.log", f"{tag}_post_check.log"), ("post_check.json", f"{tag}_post_check.json"),
                     ("bench_present.json", f"{tag}_bench_present_1bounce.json")):
        if (OUT / src).exists() and (OUT / src).stat().st_size:
            shutil.copy(OUT / src, PROF / dst)


if __name__ == "__main__":
    main()
