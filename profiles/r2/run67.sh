#!/bin/bash
# refresh of the batch-4096 ncu passes and the per-kernel times after the shared-memory fix
set -u
O=gpurun_out
T=final
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/${T}_kernel_times.jsonl 2>&1; cat $O/${T}_kernel_times.jsonl
for B in 4096; do
  timeout 100 python profiles/profile_step.py --batch $B --steps 3 > $O/${T}_plain_$B.log 2>&1 || continue
  timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/${T}_launches_b${B}.csv python profiles/profile_step.py --batch $B --steps 3 > $O/${T}_ncu_$B.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o $O/${T}_prof_b$B -f python profiles/profile_step.py --batch $B --steps 1 > $O/${T}_ncufull_$B.log 2>&1
  echo "ncu $B rc=$?"
done
