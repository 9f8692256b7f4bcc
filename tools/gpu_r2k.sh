#!/bin/bash
# Round 2, session K (1 GPU): GLSL frame shader (VRT_FRAME_GLSL) parity + timing; staging pool test; refill threshold of the trace pass.
set -x
O=gpurun_out/r2k
mkdir -p $O
timeout 600 python -m pytest tests/test_glsl_frame.py tests/test_gpu_residency.py tests/test_gpu_zz_glsl.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
VRT_EXP_N=20 timeout 300 python tools/exp.py --workload terrain --bounces 0 --flags 16 persistent=0 > $O/exp_glsl_terrain_b0.log 2>&1; tail -2 $O/exp_glsl_terrain_b0.log
VRT_EXP_N=20 timeout 300 python tools/exp.py --workload terrain --bounces 2 --flags 16 persistent=0 > $O/exp_glsl_terrain_b2.log 2>&1; tail -2 $O/exp_glsl_terrain_b2.log
VRT_EXP_N=20 timeout 300 python tools/exp.py --workload sponza --flags 16 persistent=0 > $O/exp_glsl_sponza.log 2>&1; tail -2 $O/exp_glsl_sponza.log
VRT_EXP_N=20 timeout 300 python tools/exp.py --workload sponza wavefront=1,trace_refill=24 wavefront=1,trace_refill=26 wavefront=1,trace_refill=28 wavefront=1,trace_refill=30 > $O/exp_refill_sponza.log 2>&1; tail -8 $O/exp_refill_sponza.log
ls -la $O
