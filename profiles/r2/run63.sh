#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tree" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -3
for b in 256 1024; do B2R_TICKS_PER_US=1965 B2R_LIB=profiles/micro/libb200replay_trace.so timeout 120 python profiles/micro/tree_phases.py $b; done
for e in "B2R_TREE_EARLY=1" "B2R_TREE_EARLY=0"; do
  echo "== bench sweep $e"
  env $e timeout 300 python bench.py --steps 2000 --warmup 20 --no-e2e --no-cpu-baseline | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('us/step', r['ms_per_step']*1e3, r['sweep_summary'])"
done
