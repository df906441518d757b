#!/bin/bash
set -u
run() { echo "== $*"; env "$@" timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 256,1024,4096 2>&1 | python -c "
import sys, json
for l in sys.stdin:
  try: r = json.loads(l)
  except Exception: continue
  print(r['batch'], 'fused', r['step_fused_us'], 'chain', r['chain_only_us'], 'S+G', r['sample_gather_us'])
"; }
run B2R_DEBUG_SKIP=7 B2R_GATHER_PAD_KB=0
run B2R_DEBUG_SKIP=7 B2R_NO_PRIORITY=1 B2R_GATHER_PAD_KB=0
run B2R_NO_PRIORITY=1
run B2R_NO_PRIORITY=1 B2R_GATHER_PAD_KB=0
run B2R_GATHER_PAD_KB=0
run B2R_GATHER_PAD_KB=24
