#!/bin/bash
# 1024 (copies deferred, the chain bounds the step): does throttling the copies help the chain?
set -u
run() { echo "== $*"; env "$@" timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 256,1024,2048 2>&1 | python -c "
import sys, json
for l in sys.stdin:
  try: r = json.loads(l)
  except Exception: print(l.rstrip()); continue
  print(r['batch'], 'fused', r['step_fused_us'], 'deferred', r['step_fused_deferred_us'], 'chain', r['chain_only_us'], 'gather', r['gather_us'])
"; }
run B2R_X=0
run B2R_GATHER=reg B2R_GATHER_PAD_KB=80
run B2R_GATHER=reg B2R_GATHER_PAD_KB=110
run B2R_GATHER=reg B2R_GATHER_PAD_KB=0
run B2R_GATHER=tma
