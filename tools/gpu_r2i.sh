#!/bin/bash
# Round 2, session I (1 GPU): full GPU suite, the four N=1 bench lines, edit timing (staging gather pool), launch lists and ncu --set full
# captures of the final kernels (primary frame kernel on the terrain view; wavefront passes on Sponza and on the 10 GB terrain).
set -x
O=gpurun_out/r2i
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_terrain_1gpu.json 2> $O/bench_terrain_1gpu.err; tail -2 $O/bench_terrain_1gpu.err; cut -c1-400 $O/bench_terrain_1gpu.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_1gpu.json 2> $O/bench_ref_1gpu.err; cut -c1-300 $O/bench_ref_1gpu.json
timeout 600 python bench.py --workload sponza --no-present > $O/bench_sponza_1gpu.json 2> $O/bench_sponza_1gpu.err; tail -2 $O/bench_sponza_1gpu.err; cut -c1-300 $O/bench_sponza_1gpu.json
timeout 600 python bench.py --workload edits --no-present --no-cpu > $O/bench_edits_1gpu.json 2> $O/bench_edits_1gpu.err; cut -c1-300 $O/bench_edits_1gpu.json
timeout 600 python bench.py --workload edits --edit-mode brush --no-present --no-cpu > $O/bench_edits_brush_1gpu.json 2> $O/bench_edits_brush_1gpu.err; cut -c1-300 $O/bench_edits_brush_1gpu.json
timeout 900 python bench.py --workload large --steps 10 --warmup 3 --no-present --no-cpu > $O/bench_large_1gpu.json 2> $O/bench_large_1gpu.err; tail -2 $O/bench_large_1gpu.err; cut -c1-300 $O/bench_large_1gpu.json
timeout 300 python tools/edit_timing.py 4096 1 > $O/edit_timing_1thread.log 2>&1; tail -5 $O/edit_timing_1thread.log
timeout 300 python tools/edit_timing.py 4096 0 > $O/edit_timing_pool.log 2>&1; tail -5 $O/edit_timing_pool.log
# launch lists (cold-cache, serialised: shares of the step, not absolute times)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_terrain.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-present > $O/launches_terrain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_sponza.csv python bench.py --workload sponza --wavefront 1 --steps 2 --warmup 3 --no-cpu --no-present > $O/launches_sponza.log 2>&1
# full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render -s 6 -c 1 -o $O/ncu_k_render -f python bench.py --steps 2 --warmup 3 --no-cpu --no-present > $O/ncu_k_render.log 2>&1; tail -2 $O/ncu_k_render.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 16 -c 4 -o $O/ncu_wave_sponza -f python bench.py --workload sponza --wavefront 1 --steps 2 --warmup 3 --no-cpu --no-present > $O/ncu_wave_sponza.log 2>&1; tail -2 $O/ncu_wave_sponza.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_wave -s 21 -c 7 -o $O/ncu_wave_large -f python bench.py --workload large --wavefront 1 --steps 2 --warmup 3 --no-cpu --no-present > $O/ncu_wave_large.log 2>&1; tail -2 $O/ncu_wave_large.log
ls -la $O
