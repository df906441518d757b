#!/bin/bash
# Deferred frame copies (the copies of step n beside the chain of step n + 1).
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -x -q > $O/r2_18_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_18_tests.log
timeout 300 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_18_kt.log 2>&1; cat $O/r2_18_kt.log
echo "== unsplit loss"; B2R_C51_SPLIT=0 timeout 300 python profiles/micro/kernel_times.py --per-graph 10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
  try: r = json.loads(l)
  except Exception: print(l); continue
  print(r['batch'], 'fused', r['step_fused_us'], 'deferred', r['step_fused_deferred_us'], 'chain', r['chain_only_us'])
"
