#!/bin/bash
# Round 2, session H (1 GPU): regression of the 64-bit fix and the wave-set ring, then the three bench workloads at N = 1.
set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_frames_ref.py -m gpu -x -q > $O/pytest_gpu_h.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_h.log; tail -4 $O/pytest_gpu_h.log
bash tools/gpu_r2f.sh 1
timeout 600 python bench.py --workload sponza --steps 20 --warmup 5 --no-present > $O/bench_sponza_1gpu.json 2> $O/bench_sponza_1gpu.err; cut -c1-300 $O/bench_sponza_1gpu.json
timeout 600 python bench.py --workload edits --edit-mode brush --steps 20 --warmup 5 --no-present --no-cpu > $O/bench_edits_brush_1gpu.json 2> $O/bench_edits_brush_1gpu.err; cut -c1-300 $O/bench_edits_brush_1gpu.json
