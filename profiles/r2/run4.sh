#!/bin/bash
set -u
O=gpurun_out
./profiles/micro/descent_bench > $O/r2_4_descent.log 2>&1; cat $O/r2_4_descent.log
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_4_kt.log 2>&1; cat $O/r2_4_kt.log
for B in 32 1024 4096; do
  echo "== trace new B=$B"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/trace_step.py $B 2>&1 | grep sample
done > $O/r2_4_trace.log 2>&1
cat $O/r2_4_trace.log
