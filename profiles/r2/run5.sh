#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2_5_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_5_tests.log
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_5_kt.log 2>&1; cat $O/r2_5_kt.log
B2R_SAMPLER_CLUSTER=0 timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 32 2>&1 | tail -1
for B in 32 1024; do
  echo "== trace new B=$B"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/trace_step.py $B 2>&1 | grep sample
done > $O/r2_5_trace.log 2>&1
cat $O/r2_5_trace.log
