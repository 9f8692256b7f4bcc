#!/bin/bash
# Round 2, session J (2 GPUs): flat-table slot arena (edit timing, residency tests), bounce-form tuner at N = 2.
set -x
O=gpurun_out/r2j
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_residency.py tests/test_gpu_configs.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 300 python tools/edit_timing.py 4096 0 > $O/edit_timing.log 2>&1; tail -5 $O/edit_timing.log
timeout 600 python bench.py --workload edits --no-present --no-cpu > $O/bench_edits_1gpu.json 2> $O/bench_edits_1gpu.err; cut -c1-300 $O/bench_edits_1gpu.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR bench.py --gpus 2 --workload large --steps 10 --warmup 3 --no-cpu --no-present > $O/bench_large_2gpu.json 2> $O/bench_large_2gpu.err; tail -2 $O/bench_large_2gpu.err; cut -c1-300 $O/bench_large_2gpu.json
timeout 900 $TR bench.py --gpus 2 --workload large --wavefront 1 --steps 10 --warmup 3 --no-cpu --no-present > $O/bench_large_2gpu_wave.json 2> $O/bench_large_2gpu_wave.err; cut -c1-300 $O/bench_large_2gpu_wave.json
timeout 900 $TR bench.py --gpus 2 --workload large --wavefront 0 --steps 10 --warmup 3 --no-cpu --no-present > $O/bench_large_2gpu_pixel.json 2> $O/bench_large_2gpu_pixel.err; cut -c1-300 $O/bench_large_2gpu_pixel.json
timeout 600 $TR bench.py --gpus 2 --workload edits --no-cpu --no-present > $O/bench_edits_2gpu.json 2> $O/bench_edits_2gpu.err; cut -c1-300 $O/bench_edits_2gpu.json
rm -f /dev/shm/vrt_terrain_*
ls -la $O
