#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -x -q -k "trainer" > $O/r2_44_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/r2_44_tests.log
timeout 300 python profiles/micro/e2e_breakdown.py 2>&1 | grep -v "^  " | tail -4
echo "== forked frames"; B2R_TRAINER_INLINE_FRAMES=0 timeout 300 python profiles/micro/e2e_breakdown.py 2>&1 | grep -v "^  " | tail -4
timeout 600 python bench.py --no-sweep --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('value', r['value'], 'e2e', r['e2e']['value'], r['e2e']['ms_per_step'], 'sync', r['e2e_sync']['value'])"
