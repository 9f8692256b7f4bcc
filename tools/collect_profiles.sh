#!/bin/bash
# Copies the round-2 GPU session outputs (gpurun_out/, scratch) into profiles/ (tracked) under stable names.
set -e
cd "$(dirname "$0")/.."
I=gpurun_out/r2i; F=gpurun_out/r2f; P=profiles
cpj() { [ -s "$1" ] && tail -1 "$1" > "$2" && echo "  $2"; true; }
cpj $I/bench_terrain_1gpu.json $P/r02_bench_4k_primary.json
cpj $I/bench_ref_1gpu.json $P/r02_bench_reference_arm.json
cpj $I/bench_sponza_1gpu.json $P/r02_bench_sponza_1080p_1bounce.json
cpj $I/bench_edits_1gpu.json $P/r02_bench_edits_4k.json
cpj $I/bench_edits_brush_1gpu.json $P/r02_bench_edits_brush_4k.json
cpj $I/bench_large_1gpu.json $P/r02_bench_large_10GB_4k_2bounces.json
for n in 2 4 8; do
  cpj $F/bench_terrain_${n}gpu.json $P/r02_bench_4k_primary_${n}gpu.json
  cpj $F/bench_edits_${n}gpu.json $P/r02_bench_edits_4k_${n}gpu.json
  cpj $F/bench_large_${n}gpu.json $P/r02_bench_large_10GB_4k_2bounces_${n}gpu.json
  cpj $F/bench_ref_${n}gpu.json $P/r02_bench_reference_arm_${n}gpu.json
done
[ -s $I/launches_terrain.csv ] && cp $I/launches_terrain.csv $P/r02_launches_4k_primary.csv
[ -s $I/launches_sponza.csv ] && cp $I/launches_sponza.csv $P/r02_launches_sponza_wavefront.csv
[ -s $I/pytest_gpu.log ] && cp $I/pytest_gpu.log $P/r02_pytest_gpu.txt
[ -s $I/edit_timing_pool.log ] && cat $I/edit_timing_1thread.log $I/edit_timing_pool.log > $P/r02_edit_timing.txt
ls $P | grep -c r02_
