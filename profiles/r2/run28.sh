#!/bin/bash
# TMA-staged gather against the register path at every batch size: times (CUDA events)
# and ncu --set full of both variants at 256 / 1024 / 4096 (and 32).
set -u
O=gpurun_out
for v in reg tma; do
  echo "== B2R_GATHER=$v"
  B2R_GATHER=$v timeout 300 python profiles/micro/kernel_times.py --per-graph 10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
  try: r = json.loads(l)
  except Exception: print(l.rstrip()); continue
  print(r['batch'], 'gather', r['gather_us'], 'S+G', r['sample_gather_us'], 'fused', r['step_fused_us'], 'deferred', r['step_fused_deferred_us'])
"
done
for B in 32 256 1024 4096; do
  for v in reg tma; do
    B2R_GATHER=$v timeout 100 python profiles/profile_step.py --batch $B --steps 2 > $O/plain_${v}_$B.log 2>&1 &&
    B2R_GATHER=$v timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:gather_stack4 -o $O/prof_r2_gather_${v}_b$B -f python profiles/profile_step.py --batch $B --steps 2 > $O/ncufull_${v}_$B.log 2>&1
    echo "ncu $v $B rc=$?"
  done
done
ls -la $O/*.ncu-rep
