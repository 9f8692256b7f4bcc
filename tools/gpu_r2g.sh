#!/bin/bash
# Round 2, session G (1 GPU): full GPU suite with the final kernels, the big-frame consistency check, vrt_sync timing of the edit workload.
set -x
O=gpurun_out/r2g
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 600 python tools/debug_e2e_large.py > $O/debug_e2e_large.log 2>&1; tail -16 $O/debug_e2e_large.log
timeout 300 python tools/edit_timing.py > $O/edit_timing.log 2>&1; tail -30 $O/edit_timing.log
ls $O
