#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2_10_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_10_tests.log
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_10_kt.log 2>&1; cat $O/r2_10_kt.log
B2R_ROW_FLAGS=0 timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 256,1024,4096 2>&1 | tail -3
B2R_GATHER_PAD_KB=0 timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 256,1024 2>&1 | tail -2
for B in 1024 4096; do
  echo "== timeline B=$B"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py $B 2>&1 | tail -32
done > $O/r2_10_timeline.log 2>&1
cat $O/r2_10_timeline.log
