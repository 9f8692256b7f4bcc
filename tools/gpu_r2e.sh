#!/bin/bash
# Round 2, session E (2 GPUs): where the N > 1 gather costs its ~9 % (owner variants, payload size, stream count) + trace pass v3 on one GPU.
set -x
O=gpurun_out/r2e
mkdir -p $O
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 400 python tools/exp.py --workload terrain --bounces 1 wavefront=1 > $O/exp_terrain_b1.log 2>&1; tail -2 $O/exp_terrain_b1.log
timeout 400 python tools/exp.py --workload sponza wavefront=1 > $O/exp_sponza.log 2>&1; tail -2 $O/exp_sponza.log
run() { tag=$1; shift; timeout 300 env "$@" $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu $EXTRA > $O/bench_$tag.json 2> $O/bench_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$tag.json").read().strip().splitlines()[-1]); g=d["gather"]
    print("RESULT $tag", round(d["value"]), "trace-only", round(g["value_trace_only_warm_l2"]), "lat", round(g["frame_latency_ms"],3), "e2e", round(d["e2e"]["value"]))
except Exception as e: print("RESULT $tag ERR", e)
PY
}
EXTRA=""
run rotate VRT_BENCH_OWNER=rotate
run self VRT_BENCH_OWNER=self
run zero VRT_BENCH_OWNER=zero
run streams2 VRT_BENCH_STREAMS=2
run streams1 VRT_BENCH_STREAMS=1
EXTRA="--no-compact"
run rotate_full VRT_BENCH_OWNER=rotate
EXTRA="--steps 100"
run rotate_100 VRT_BENCH_OWNER=rotate
ls $O
