#!/bin/bash
# write-back without a hand-over behind the values: parity, timelines, sweep
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tree" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -x -q -k "1024 or 256 or shard" 2>&1 | tail -5
for b in 256 1024; do
  echo "== timeline batch $b defer"
  B2R_LIB=profiles/micro/libb200replay_trace.so timeout 200 python profiles/micro/timeline.py $b 1000000 defer 2>&1 | grep -E "S END|L-tail loss|T apply|T leaf|level |write-back"
done
for e in 1 0; do
  echo "== bench sweep B2R_TREE_EARLY=$e"
  B2R_TREE_EARLY=$e timeout 300 python bench.py --steps 2000 --warmup 20 --no-e2e --no-cpu-baseline | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('us/step', r['ms_per_step']*1e3, r['sweep_summary'])"
done
