#!/bin/bash
# The driver's SCALE sequence at N=2: complete default bench (e2e, sweeps, DDP step), both arms.
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_36_n2.json 2> $O/r2_36_n2.err; echo "n2 rc=$?"; python - <<'PY'
import json
r = json.load(open('gpurun_out/r2_36_n2.json'))
for k in ('value', 'ms_per_step', 'sweep_summary', 'e2e', 'shard_check', 'timing', 'full_train_step'):
  print(k, str(r.get(k))[:500])
PY
tail -3 $O/r2_36_n2.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/r2_36_ref_n2.json 2> $O/r2_36_ref.err; echo "ref rc=$?"; head -c 600 $O/r2_36_ref_n2.json
