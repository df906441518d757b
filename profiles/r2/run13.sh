#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2_13_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_13_tests.log
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_13_kt.log 2>&1; cat $O/r2_13_kt.log
B2R_FUSE_WRITEBACK=0 timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 32 2>&1 | tail -1
echo "== timeline B=32"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py 32 2>&1 | tail -32
