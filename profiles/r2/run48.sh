#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_loss_goldens.py -m gpu -x -q > $O/r2_48_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_48_tests.log
for i in 1 2 3; do timeout 300 python profiles/micro/kernel_times.py --per-graph 10 --batches 32 2>&1 | tail -1 | python -c "
import sys, json
r=json.loads(sys.stdin.read()); print('fused', r['step_fused_us'], 'chain', r['chain_only_us'], 'sample', r['sample_us'])"; done
timeout 300 python profiles/micro/kernel_times.py --per-graph 10 --batches 256,1024 2>&1 | tail -2 | python -c "
import sys, json
for l in sys.stdin:
  r=json.loads(l); print(r['batch'], 'fused', r['step_fused_us'], 'deferred', r['step_fused_deferred_us'], 'chain', r['chain_only_us'])"
