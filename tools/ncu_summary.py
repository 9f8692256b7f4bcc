#!/usr/bin/env python
"""ncu report -> tracked summaries under profiles/ and an entry of profiles/ncu_summary.json (what bench.py's roofline block cites).

    python tools/ncu_summary.py <report.ncu-rep> <name> [--key "terrain 3840x2160 bounces=0 gpus=1"] [--regions <libvoxelrt_b200.so> <csrc dir>]

Writes profiles/<name>_details.txt (ncu --page details), profiles/<name>_raw_summary.csv (selected metrics, one column per captured launch),
optionally profiles/<name>_source_regions.txt (tools/regions.py), and — with --key — the per-launch numbers of every captured kernel under that
key in profiles/ncu_summary.json.
"""
import csv
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PROF = ROOT / "profiles"
METRICS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum lts__t_bytes.sum lts__t_sectors.sum l1tex__t_sector_hit_rate.pct
lts__t_sector_hit_rate.pct smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__grid_size launch__block_size
lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio""".split()
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def num(v):
    return float(v.replace(",", ""))


def main():
    rep, name = Path(sys.argv[1]), sys.argv[2]
    key = sys.argv[sys.argv.index("--key") + 1] if "--key" in sys.argv else None
    det = subprocess.run(["ncu", "-i", str(rep), "--page", "details"], capture_output=True, text=True).stdout
    (PROF / f"{name}_details.txt").write_text(det)
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, launches = rows[0], rows[1], rows[2:]
    kn = head.index("Kernel Name")
    with open(PROF / f"{name}_raw_summary.csv", "w") as f:
        f.write("metric,unit," + ",".join(f"launch{i}" for i in range(len(launches))) + "\n")
        f.write("kernel,," + ",".join('"' + r[kn].replace('"', "'") + '"' for r in launches) + "\n")
        for m in METRICS:
            if m in head:
                i = head.index(m)
                f.write(f"{m},{units[i]}," + ",".join(r[i].replace(",", "") for r in launches) + "\n")
    if "--regions" in sys.argv:
        k = sys.argv.index("--regions")
        src_csv = PROF / f".{name}_source.csv"
        src_csv.write_text(subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout)
        env = dict(os.environ, VRT_REGIONS_LIB=sys.argv[k + 1], VRT_REGIONS_SRC=sys.argv[k + 2])
        subprocess.run([sys.executable, str(ROOT / "tools" / "regions.py"), str(src_csv), str(PROF / f"{name}_source_regions.txt")], stdout=subprocess.DEVNULL, env=env)
        src_csv.unlink()
    if key:
        def get(r, m, scale=False):
            if m not in head:
                return None
            i = head.index(m)
            return num(r[i]) * (SCALE.get(units[i], 1.0) if scale else 1.0)

        kernels = []
        for r in launches:
            kernels.append({
                "kernel": r[kn].split("(")[0].replace("void ", ""),
                "duration_us": round(get(r, "gpu__time_duration.sum", True) * 1e6, 2),
                "dram_bytes": int(get(r, "dram__bytes_read.sum", True) + get(r, "dram__bytes_write.sum", True)),
                # L2 traffic: lts__t_bytes where the set carries it, else 32-byte sectors
                "l2_bytes": int(get(r, "lts__t_bytes.sum", True)) if get(r, "lts__t_bytes.sum") is not None
                else (None if get(r, "lts__t_sectors.sum") is None else int(get(r, "lts__t_sectors.sum")) * 32),
                "warp_instructions": int(get(r, "smsp__inst_executed.sum")),
                "threads_per_instruction": get(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                "issue_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "alu_pipe_pct": get(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "fma_pipe_pct": get(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "l1_hit_pct": get(r, "l1tex__t_sector_hit_rate.pct"),
                "l2_hit_pct": get(r, "lts__t_sector_hit_rate.pct"),
                "registers": int(get(r, "launch__registers_per_thread")),
            })
        path = PROF / "ncu_summary.json"
        d = json.loads(path.read_text()) if path.exists() else {"_comment": "per-launch ncu numbers (ncu --set full --clock-control none; cold-cache, serialised launches) of the kernels one bench step launches, by workload; written by tools/ncu_summary.py, cited by bench.py's roofline block"}
        d[key] = {"capture": f"profiles/{name}_raw_summary.csv", "dram_bytes": sum(k["dram_bytes"] for k in kernels), "kernels": kernels}
        path.write_text(json.dumps(d, indent=1) + "\n")
    print("wrote", name)


if __name__ == "__main__":
    main()
