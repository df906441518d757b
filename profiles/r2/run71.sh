#!/bin/bash
# 1024 rows, deferred copies: the copies' occupancy cap against today's (shorter) chain
set -u
for m in 512 1024; do for pad in 0 28 36 44; do
  printf "batch 1024 B2R_TREE_EARLY_MAX=$m B2R_GATHER_PAD_KB=$pad: "
  B2R_TREE_EARLY_MAX=$m B2R_GATHER_PAD_KB=$pad timeout 200 python bench.py --batch 1024 --steps 2000 --warmup 20 --no-e2e --no-cpu-baseline --no-sweep | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('us/step %.2f' % (r['ms_per_step']*1e3))"
done; done
