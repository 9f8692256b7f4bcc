#!/bin/bash
# Round 2, session F (N GPUs): BASELINE configs[1] scaling point, configs[4] (edits) and configs[3] (10 GB terrain, 2 bounces) at N GPUs,
# reference arm under torchrun.
set -x
N=${1:-8}
O=gpurun_out/r2f
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
[ "$N" = "1" ] && TR="python"
nvidia-smi -L | head -8 > $O/gpus_${N}.txt; nproc >> $O/gpus_${N}.txt; free -g | head -2 >> $O/gpus_${N}.txt; df -h /dev/shm | tail -1 >> $O/gpus_${N}.txt
summ() { python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1]); g=d.get("gather") or {}
    print("RESULT $1", "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "host_ok", d["e2e"].get("host_frame_equal_to_single_gpu_frame"),
          "trace-only", g.get("value_trace_only_warm_l2") and round(g["value_trace_only_warm_l2"]), "gather_ok", g.get("verified_equal_to_single_gpu_frame"), "replicas", g.get("replicas_identical"),
          "frac", round(d["roofline"]["frac"],3), "upload", d["residency"].get("scene_upload_GB_per_s"))
except Exception as e: print("RESULT $1 ERR", e)
PY
}
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-present > $O/bench_terrain_${N}gpu.json 2> $O/bench_terrain_${N}gpu.err; summ $O/bench_terrain_${N}gpu.json
timeout 600 $TR bench.py --gpus $N --workload edits --steps 20 --warmup 5 --no-cpu --no-present > $O/bench_edits_${N}gpu.json 2> $O/bench_edits_${N}gpu.err; summ $O/bench_edits_${N}gpu.json
timeout 900 $TR bench.py --gpus $N --workload large --steps 10 --warmup 3 --no-cpu --no-present > $O/bench_large_${N}gpu.json 2> $O/bench_large_${N}gpu.err; summ $O/bench_large_${N}gpu.json; tail -3 $O/bench_large_${N}gpu.err
if [ "$N" != "1" ]; then timeout 300 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/bench_ref_${N}gpu.json 2> $O/bench_ref_${N}gpu.err; cut -c1-200 $O/bench_ref_${N}gpu.json; fi
rm -f /dev/shm/vrt_terrain_*
if [ "$N" != "1" ]; then timeout 300 $TR tools/d2h_ceiling.py > $O/d2h_ceiling_${N}gpu.txt 2> $O/d2h_ceiling_${N}gpu.err; cat $O/d2h_ceiling_${N}gpu.txt; tail -3 $O/d2h_ceiling_${N}gpu.err; fi
ls -la $O
