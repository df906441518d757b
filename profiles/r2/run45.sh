#!/bin/bash
# Cluster sampler: closing by warp 0 alone when no pick is invalid.
set -u
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -m gpu -x -q > $O/r2_45_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_45_tests.log
timeout 300 python profiles/micro/kernel_times.py --per-graph 10 --batches 32 2>&1 | tail -1
echo "== timeline"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py 32 2>&1 | tail -22
