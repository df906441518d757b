#!/bin/bash
# Round-2 run 1: parity suite on the new sampler / tiny tree kernels, per-kernel times
# new vs old, C51 error survey.
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2_1_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_1_tests.log
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_1_kt_new.log 2>&1; cat $O/r2_1_kt_new.log
B2R_SAMPLER=thread B2R_TREE_TINY=0 timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_1_kt_old.log 2>&1; cat $O/r2_1_kt_old.log
timeout 200 python profiles/micro/c51_error.py > $O/r2_1_c51err.log 2>&1; cat $O/r2_1_c51err.log
