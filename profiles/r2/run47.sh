#!/bin/bash
# Cluster sampler: flags / speculative draws / minima handed to CTA 0 by st.async onto an mbarrier.
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py tests/test_actor.py -m gpu -x -q > $O/r2_47_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_47_tests.log
timeout 300 python profiles/micro/kernel_times.py --per-graph 10 --batches 32 2>&1 | tail -1
echo "== timeline"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py 32 2>&1 | tail -18
