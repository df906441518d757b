#!/bin/bash
# Sharded step: window range finding, device-side tiny/one-CTA choice, deferred copies.
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -m gpu -x -q > $O/r2_23_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_23_tests.log
for w in 2 8; do
echo "== sharded timeline per-GPU 32, world $w"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 200 python profiles/micro/timeline_sharded.py 32 $w 2>&1 | tail -32
done
