#!/bin/bash
# Round 2, session D (2 GPUs): the N > 1 paths — compact gather, shared host frame for e2e, edits with replica digest, big scene shared over
# /dev/shm, reference arm under torchrun — and the tile-order A/B of the primary kernel on one GPU.
set -x
O=gpurun_out/r2d
mkdir -p $O
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 python tools/exp.py tile_order=0 tile_order=1 > $O/exp_tile_order.log 2>&1; tail -4 $O/exp_tile_order.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_terrain_${N}gpu.json 2> $O/bench_terrain_${N}gpu.err; tail -3 $O/bench_terrain_${N}gpu.err; cut -c1-400 $O/bench_terrain_${N}gpu.json
timeout 300 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/bench_ref_${N}gpu.json 2> $O/bench_ref_${N}gpu.err; cut -c1-300 $O/bench_ref_${N}gpu.json
timeout 600 $TR bench.py --gpus $N --workload edits --steps 20 --warmup 5 --no-cpu > $O/bench_edits_${N}gpu.json 2> $O/bench_edits_${N}gpu.err; tail -3 $O/bench_edits_${N}gpu.err; cut -c1-300 $O/bench_edits_${N}gpu.json
timeout 900 $TR bench.py --gpus $N --workload large --large-y 7 --large-shift 0 --steps 10 --warmup 3 --no-cpu > $O/bench_large_small_${N}gpu.json 2> $O/bench_large_small_${N}gpu.err; tail -3 $O/bench_large_small_${N}gpu.err; cut -c1-300 $O/bench_large_small_${N}gpu.json
ls -la $O; ls /dev/shm | head
