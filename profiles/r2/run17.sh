#!/bin/bash
# Loss in two halves, v2: PreSync hand-shake, multi-CTA tail, stats from the first half.
set -u
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py tests/test_loss_goldens.py tests/test_learner.py -m gpu -x -q > $O/r2_17_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2_17_tests.log
timeout 300 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_17_kt.log 2>&1; cat $O/r2_17_kt.log
for b in 32 256 1024 4096; do
echo "== timeline B=$b split"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py $b 2>&1 | tail -40
done
