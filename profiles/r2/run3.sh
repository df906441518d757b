#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2_3_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_3_tests.log
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_3_kt_new.log 2>&1; cat $O/r2_3_kt_new.log
for B in 32 1024 4096; do
  echo "== trace new B=$B"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/trace_step.py $B 2>&1 | grep sample
done > $O/r2_3_trace.log 2>&1
cat $O/r2_3_trace.log
B2R_SAMPLER_WAVE_CTAS=100000 timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 4096 2>&1 | tail -1
