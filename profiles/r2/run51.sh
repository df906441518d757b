#!/bin/bash
# write-back timelines inside the fused step
set -u
for b in 256 1024 4096; do
  echo "== timeline batch $b"
  B2R_LIB=profiles/micro/libb200replay_trace.so timeout 200 python profiles/micro/timeline.py $b 2>&1 | tail -45
done
