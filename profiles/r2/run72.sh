#!/bin/bash
# 513..1536 rows: the early write-back with the copies unrestricted (register kernel without
# its shared-memory pad, or the TMA kernel)
set -u
run() {
  printf "batch $1 $2: "
  env $2 timeout 200 python bench.py --batch $1 --steps 2000 --warmup 20 --no-e2e --no-cpu-baseline --no-sweep | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('us/step %.2f' % (r['ms_per_step']*1e3))"
}
run 1024 "B2R_TREE_EARLY_MAX=1024 B2R_GATHER=tma"
run 768 "B2R_X=0"
run 768 "B2R_TREE_EARLY_MAX=1024 B2R_GATHER_PAD_KB=0"
run 768 "B2R_TREE_EARLY_MAX=1024 B2R_GATHER=tma"
run 1536 "B2R_X=0"
run 1536 "B2R_GATHER_PAD_KB=0"
