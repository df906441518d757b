#!/bin/bash
# Round-1 closing measurements on one B200 (run through gpurun from the repo root):
# tests, smoke, both bench arms, per-kernel times, then the ncu passes (launch lists
# and one full capture per batch size) of commands that have just exited 0 without ncu.
set -u
O=gpurun_out
T=${1:-v4}
timeout 400 python -m pytest tests -m gpu -x -q > $O/final_${T}_tests.log 2>&1; echo "rc=$?" >> $O/final_${T}_tests.log
tail -2 $O/final_${T}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/final_${T}_smoke.log 2>&1; tail -1 $O/final_${T}_smoke.log
timeout 400 python bench.py > $O/final_${T}_n1.json 2> $O/final_${T}_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > $O/final_${T}_ref.json 2> $O/final_${T}_ref.err; echo "ref rc=$?"
timeout 100 python profiles/micro/kernel_times.py --per-graph 10 > $O/kernel_times_${T}.log 2>&1; cat $O/kernel_times_${T}.log
[ "${2:-}" = "no-ncu" ] && exit 0
for B in 32 4096; do
  timeout 100 python profiles/profile_step.py --batch $B --steps 3 > $O/plain_${T}_$B.log 2>&1 || continue
  timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/launches_b${B}_${T}.csv python profiles/profile_step.py --batch $B --steps 3 > $O/ncu_${T}_$B.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o $O/prof_${T}_b$B -f python profiles/profile_step.py --batch $B --steps 1 > $O/ncufull_${T}_$B.log 2>&1
done
ls -la $O | tail -15
