#!/bin/bash
set -u
O=gpurun_out
for B in 32 256 1024 4096; do
  echo "== timeline B=$B"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py $B 2>&1 | tail -32
done > $O/r2_8_timeline.log 2>&1
cat $O/r2_8_timeline.log
