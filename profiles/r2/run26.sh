#!/bin/bash
# Host-batch path: breakdown of one update.
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_checkpoint.py tests/test_actor.py -m gpu -x -q > $O/r2_26_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_26_tests.log
timeout 300 python profiles/micro/host_batch_breakdown.py 32 2>&1 | tail -12
