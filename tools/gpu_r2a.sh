#!/bin/bash
# Round 2, session A (1 GPU): parity suite with the packet-exact frame kernels and the new wavefront pipeline, A/B timings, ncu captures.
set -x
mkdir -p gpurun_out
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi -L > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 300 python tools/exp.py "persistent=0" > $O/exp_primary.log 2>&1; tail -2 $O/exp_primary.log
timeout 400 python tools/exp.py --workload terrain --bounces 1 wavefront=0 wavefront=1 > $O/exp_terrain_b1.log 2>&1; tail -4 $O/exp_terrain_b1.log
timeout 400 python tools/exp.py --workload terrain --bounces 2 wavefront=0 wavefront=1 > $O/exp_terrain_b2.log 2>&1; tail -4 $O/exp_terrain_b2.log
timeout 400 python tools/exp.py --workload sponza wavefront=0 wavefront=1 > $O/exp_sponza.log 2>&1; tail -4 $O/exp_sponza.log
timeout 600 python tools/exp.py --workload large wavefront=0 wavefront=1 > $O/exp_large.log 2>&1; tail -4 $O/exp_large.log
export VRT_EXP_N=2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_terrain_b1_wave.csv python tools/exp.py --workload terrain --bounces 1 wavefront=1 > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wave_trace -s 3 -c 1 -f -o $O/prof_wave_trace python tools/exp.py --workload terrain --bounces 1 wavefront=1 > $O/ncu_wave_trace.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_wave_shade -s 3 -c 1 -f -o $O/prof_wave_shade python tools/exp.py --workload terrain --bounces 1 wavefront=1 > $O/ncu_wave_shade.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render -s 3 -c 1 -f -o $O/prof_render python tools/exp.py "persistent=0" > $O/ncu_render.log 2>&1
ls -la $O
