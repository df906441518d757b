#!/bin/bash
set -u
for skip in 0 1 2 3 4 7; do
  echo "== skip=$skip (1 loss, 2 write-back, 4 presort)"
  B2R_DEBUG_SKIP=$skip timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 1024,4096 2>&1 | python -c "
import sys, json
for l in sys.stdin:
  try: r = json.loads(l)
  except Exception: continue
  print(r['batch'], 'fused', r['step_fused_us'], 'chain', r['chain_only_us'], 'S+G', r['sample_gather_us'])
"
done
echo "== skip=0 no row flags"; B2R_ROW_FLAGS=0 B2R_DEBUG_SKIP=7 timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 1024,4096 2>&1 | tail -2 | cut -c1-220
