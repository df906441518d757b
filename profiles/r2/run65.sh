#!/bin/bash
# N GPUs of one box: the sharded step (value, sweep, e2e) and the reference arm's line
set -u
N=${1:-8}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 500 $TR bench.py --gpus $N --steps 4000 --warmup 20 > $O/bench_n${N}_run65.out 2> $O/bench_n${N}_run65.err; echo "rc=$?"
grep '^{' $O/bench_n${N}_run65.out | tail -1 > $O/bench_n${N}_run65.json
python - <<PY
import json
r = json.load(open('$O/bench_n${N}_run65.json'))
print('N=$N us/step', r['ms_per_step'] * 1e3, 'value', r['value'], 'e2e', r.get('e2e', {}).get('value'),
      r.get('shard_check', {}).get('rows_per_rank'), r.get('sweep_summary'))
PY
tail -3 $O/bench_n${N}_run65.err
