#!/bin/bash
set -u
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2_38_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/r2_38_tests.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print(r['value'], r['ms_per_step'], r['sweep_summary'])"
