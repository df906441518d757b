#!/bin/bash
# new defaults (early write-back up to 1024 rows, TMA copies at every size): parity and sweep
set -u
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 2000 --warmup 20 --no-e2e --no-cpu-baseline | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('us/step', r['ms_per_step']*1e3, r['roofline']['kernel'], r['sweep_summary'])"
