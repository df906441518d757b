#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2_6_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_6_tests.log
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_6_kt.log 2>&1; cat $O/r2_6_kt.log
B2R_TREE_PRESORT=0 timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 256,1024,4096 2>&1 | tail -3
