#!/bin/bash
# Round 2, session Q (1 GPU): bench lines of the GLSL frame shader (--glsl: VRT_FRAME_GLSL frames, rays counted like GpuRenderer.cpp:292-296).
set -x
O=gpurun_out/r2q
mkdir -p $O
timeout 300 python bench.py --glsl --no-cpu --no-present > $O/bench_glsl_terrain_b0.json 2> $O/bench_glsl_terrain_b0.err; cut -c1-260 $O/bench_glsl_terrain_b0.json; tail -2 $O/bench_glsl_terrain_b0.err
timeout 300 python bench.py --glsl --bounces 2 --steps 10 --no-cpu --no-present > $O/bench_glsl_terrain_b2.json 2> $O/bench_glsl_terrain_b2.err; cut -c1-260 $O/bench_glsl_terrain_b2.json; tail -2 $O/bench_glsl_terrain_b2.err
timeout 300 python bench.py --glsl --workload sponza --no-cpu --no-present > $O/bench_glsl_sponza.json 2> $O/bench_glsl_sponza.err; cut -c1-260 $O/bench_glsl_sponza.json; tail -2 $O/bench_glsl_sponza.err
