#!/bin/bash
# Deferred frame copies: how many copy CTAs per SM leave room for the next step's chain.
set -u
run() { echo "== $*"; env "$@" timeout 300 python profiles/micro/kernel_times.py --per-graph 10 --batches 256,1024,4096 2>&1 | python -c "
import sys, json
for l in sys.stdin:
  try: r = json.loads(l)
  except Exception: print(l.rstrip()); continue
  print(r['batch'], 'fused', r['step_fused_us'], 'deferred', r['step_fused_deferred_us'], 'chain', r['chain_only_us'], 'S+G', r['sample_gather_us'])
"; }
run B2R_X=0
run B2R_GATHER_PAD_KB=28
run B2R_GATHER_PAD_KB=56
run B2R_GATHER_PAD_KB=28 B2R_C51_SPLIT=0
run B2R_GATHER_PAD_KB=18
