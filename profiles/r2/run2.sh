#!/bin/bash
set -u
O=gpurun_out
for B in 32 1024 4096; do
  echo "== trace new B=$B"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/trace_step.py $B 2>&1 | tail -5
  echo "== trace old B=$B"; B2R_SAMPLER=thread B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/trace_step.py $B 2>&1 | tail -5
done > $O/r2_2_trace.log 2>&1
cat $O/r2_2_trace.log
