#!/bin/bash
# One GPU session: parity tests, bench, ncu launch list, ncu full capture of the render kernel.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
grep -m1 -o "avx512f" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_post_check.py > /dev/null 2>&1; tail -4 gpurun_out/post_check.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 python bench.py --bounces 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_present.json 2> gpurun_out/bench_present.err; tail -1 gpurun_out/bench_present.json | cut -c1-400
if [ "$1" != "noprof" ]; then
timeout 600 ncu --set full --clock-control none -k regex:k_atrous -s 2 -c 5 -f -o gpurun_out/prof_atrous python bench.py --bounces 1 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_atrous.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render -s 4 -c 2 -f -o gpurun_out/prof_render python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_render.ncu-rep --page source --csv > gpurun_out/prof_render_source.csv 2>/dev/null
# wavefront passes of a 2-bounce frame (level 1: budgets 16 / 32 / rest); the per-pixel bounce kernel's capture is r01_ncu_k_render_bounce1_*
timeout 600 ncu --set full --clock-control none -k regex:k_wave_trace -s 6 -c 3 -f -o gpurun_out/prof_wave python tools/exp.py --workload terrain --bounces 2 wavefront=1 > gpurun_out/ncu_wave.log 2>&1
timeout 300 python bench.py --workload edits --edit-mode brush --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_edits_brush.json 2>/dev/null
for wl in sponza large edits; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -1 gpurun_out/bench_$wl.json | cut -c1-300
done
fi
ls -la gpurun_out
