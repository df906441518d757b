#!/bin/bash
set -u
for mode in "" defer; do
for b in 256 1024; do
  echo "== timeline batch $b $mode"
  B2R_LIB=profiles/micro/libb200replay_trace.so timeout 200 python profiles/micro/timeline.py $b 1000000 $mode 2>&1 | grep -E "S END|L-tail loss|T apply|T leaf|level |write-back"
done
done
