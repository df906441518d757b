#!/bin/bash
# Suite after the loss split / max_rows / tolerance changes; timelines of the step.
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_16_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2_16_tests.log
for b in 32 1024; do
echo "== timeline B=$b split"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py $b 2>&1 | tail -40
echo "== timeline B=$b unsplit"; B2R_C51_SPLIT=0 B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py $b 2>&1 | tail -40
done
