#!/bin/bash
# 1024 rows, deferred copies: are the copies held back by the chain's launch priority?
set -u
for e in "B2R_X=0" "B2R_NO_PRIORITY=1" "B2R_GATHER=tma" "B2R_NO_L2_PERSIST=1"; do
  printf "batch 1024 $e: "
  env $e timeout 200 python bench.py --batch 1024 --steps 2000 --warmup 20 --no-e2e --no-cpu-baseline --no-sweep | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('us/step %.2f' % (r['ms_per_step']*1e3))"
done
