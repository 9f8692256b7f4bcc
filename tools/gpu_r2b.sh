#!/bin/bash
# Round 2, session B (1 GPU): trace kernel v2 (classification at the producer, generic rays apart, pinned constants) and the primary kernel's
# single-vote packet test / constant-memory rsqrt14 table.
set -x
O=gpurun_out/r2b
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_frames_ref.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 300 python tools/exp.py "persistent=0" > $O/exp_primary.log 2>&1; tail -2 $O/exp_primary.log
timeout 600 python tools/exp.py --workload terrain --bounces 1 wavefront=0 wavefront=1,trace_refill=24 wavefront=1,trace_refill=16 wavefront=1,trace_refill=20 wavefront=1,trace_refill=28 wavefront=1,trace_refill=31 > $O/exp_terrain_b1.log 2>&1; tail -6 $O/exp_terrain_b1.log
timeout 400 python tools/exp.py --workload sponza wavefront=0 wavefront=1,trace_refill=24 wavefront=1,trace_refill=28 > $O/exp_sponza.log 2>&1; tail -3 $O/exp_sponza.log
timeout 600 python tools/exp.py --workload large wavefront=0 wavefront=1,trace_refill=24 wavefront=1,trace_refill=28 > $O/exp_large.log 2>&1; tail -3 $O/exp_large.log
export VRT_EXP_N=2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave_trace -s 6 -c 2 -f -o $O/prof_wave_trace python tools/exp.py --workload terrain --bounces 1 wavefront=1 > $O/ncu_wave_trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave_primary -s 3 -c 1 -f -o $O/prof_wave_primary python tools/exp.py --workload terrain --bounces 1 wavefront=1 > $O/ncu_wave_primary.log 2>&1
ls -la $O
