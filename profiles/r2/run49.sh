#!/bin/bash
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --steps 4000 --warmup 20 --no-e2e 2>$O/r2_49.err | python -c "
import sys, json
for l in sys.stdin:
  try: r = json.loads(l)
  except Exception: continue
  print('N=2 us/step', r['ms_per_step']*1e3, 'value', r['value'], r.get('shard_check',{}).get('rows_per_rank'), r['sweep_summary'])
"
tail -2 $O/r2_49.err
timeout 300 python bench.py --steps 4000 --warmup 20 --no-e2e --no-cpu-baseline --no-sweep | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('N=1 us/step', r['ms_per_step']*1e3)"
echo "== sharded timeline world 2"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 200 python profiles/micro/timeline_sharded.py 32 2 2>&1 | tail -24
