#!/bin/bash
set -u
for b in 64 128 512; do for e in 1024 0; do
  printf "batch $b B2R_TREE_EARLY_MAX=$e: "
  B2R_TREE_EARLY_MAX=$e timeout 200 python bench.py --batch $b --steps 2000 --warmup 20 --no-e2e --no-cpu-baseline --no-sweep | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('us/step %.2f' % (r['ms_per_step']*1e3))"
done; done
