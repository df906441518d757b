#!/bin/bash
# One shard's step on one GPU (ranks emulated): where the sharded step loses against the plain one.
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -m gpu -x -q -k "shard or tree" > $O/r2_22_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_22_tests.log
for w in 2 8; do
echo "== sharded timeline per-GPU 32, world $w"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 200 python profiles/micro/timeline_sharded.py 32 $w 2>&1 | tail -40
done
echo "== plain timeline 32"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py 32 2>&1 | tail -30
