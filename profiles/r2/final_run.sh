#!/bin/bash
# Round-2 closing measurements on one B200 (run through gpurun from the repo root):
# tests, smoke, both bench arms, per-kernel times, then the ncu passes (launch lists and
# one full capture per batch size) of commands that have just exited 0 without ncu.
set -u
O=gpurun_out
T=${1:-final}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "rc=$?" >> $O/${T}_tests.log
tail -2 $O/${T}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
timeout 600 python bench.py > $O/${T}_n1.json 2> $O/${T}_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > $O/${T}_n1_k20.json 2> $O/${T}_n1_k20.err; echo "bench k20 rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > $O/${T}_ref.json 2> $O/${T}_ref.err; echo "ref rc=$?"
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/${T}_kernel_times.jsonl 2>&1; cat $O/${T}_kernel_times.jsonl
[ "${2:-}" = "no-ncu" ] && exit 0
for B in 32 256 4096; do
  timeout 100 python profiles/profile_step.py --batch $B --steps 3 > $O/${T}_plain_$B.log 2>&1 || continue
  timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/${T}_launches_b${B}.csv python profiles/profile_step.py --batch $B --steps 3 > $O/${T}_ncu_$B.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o $O/${T}_prof_b$B -f python profiles/profile_step.py --batch $B --steps 1 > $O/${T}_ncufull_$B.log 2>&1
  echo "ncu $B rc=$?"
done
ls -la $O | grep "${T}_" | tail -20
