#!/bin/bash
set -u
run() {
  printf "batch $1 $2: "
  env $2 timeout 200 python bench.py --batch $1 --steps 2000 --warmup 20 --no-e2e --no-cpu-baseline --no-sweep | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('us/step %.2f' % (r['ms_per_step']*1e3))"
}
run 1536 "B2R_GATHER=tma"
run 2048 "B2R_X=0"
run 2048 "B2R_GATHER=reg B2R_GATHER_PAD_KB=0"
