#!/bin/bash
# GPU session for the non-headline BASELINE configs: parity tests + one bench line per workload.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/pytest_configs.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_configs.log
tail -5 gpurun_out/pytest_configs.log
for wl in sponza large edits; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -2 gpurun_out/bench_$wl.err; cat gpurun_out/bench_$wl.json
done
timeout 300 python bench.py --bounces 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_terrain_b1.json 2>/dev/null; cat gpurun_out/bench_terrain_b1.json
