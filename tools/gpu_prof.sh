#!/bin/bash
# GPU session: parity tests, macro statistics, bench, ncu full capture (source counters) of the render kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/macro_stats.py > gpurun_out/macro_stats.log 2>&1; cat gpurun_out/macro_stats.log | head -20
timeout 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" != "noprof" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render -s 4 -c 1 -o gpurun_out/prof_render -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_render.ncu-rep --page source --csv > gpurun_out/prof_render_source.csv 2>/dev/null
fi
ls -la gpurun_out
