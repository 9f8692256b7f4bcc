#!/bin/bash
# Round 2, session L (4 GPUs): the edit workload at N = 4 with the flat-table slot arena (the N = 4 line of session F predates it).
set -x
O=gpurun_out/r2l
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519"
timeout 400 $TR bench.py --gpus 4 --workload edits --steps 20 --warmup 5 --no-cpu --no-present > $O/bench_edits_4gpu.json 2> $O/bench_edits_4gpu.err; cut -c1-300 $O/bench_edits_4gpu.json; tail -2 $O/bench_edits_4gpu.err
