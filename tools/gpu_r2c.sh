#!/bin/bash
# Round 2, session C (1 GPU): occupancy of the trace pass (8 / 10 / 12 CTAs per SM), generic rays on a side stream; bench.py with the compact payload.
set -x
O=gpurun_out/r2c
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_frames_ref.py tests/test_gpu_parity.py tests/test_gpu_residency.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 600 python tools/exp.py --workload terrain --bounces 1 wavefront=0 wavefront=1,trace_ctas=8 wavefront=1,trace_ctas=10 wavefront=1,trace_ctas=12 wavefront=1,trace_ctas=10,trace_refill=20 wavefront=1,trace_ctas=12,trace_refill=20 > $O/exp_terrain_b1.log 2>&1; tail -6 $O/exp_terrain_b1.log
timeout 400 python tools/exp.py --workload sponza wavefront=0 wavefront=1,trace_ctas=8,trace_refill=24 wavefront=1,trace_ctas=10 wavefront=1,trace_ctas=12 > $O/exp_sponza.log 2>&1; tail -4 $O/exp_sponza.log
timeout 600 python tools/exp.py --workload large wavefront=1,trace_ctas=8 wavefront=1,trace_ctas=10 wavefront=1,trace_ctas=12 > $O/exp_large.log 2>&1; tail -3 $O/exp_large.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-present > $O/bench_terrain.json 2> $O/bench_terrain.err; tail -2 $O/bench_terrain.err; cut -c1-600 $O/bench_terrain.json
timeout 600 python bench.py --workload sponza --steps 10 --warmup 3 --no-present --no-cpu > $O/bench_sponza.json 2> $O/bench_sponza.err; tail -2 $O/bench_sponza.err; cut -c1-300 $O/bench_sponza.json
ls -la $O
