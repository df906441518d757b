#!/bin/bash
# Shard's step: write-back inside the loss tail's cluster (device-side row count).
set -u
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -m gpu -x -q > $O/r2_35_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/r2_35_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
for flag in "" "--no-early-publish"; do
  echo "== N=2 $flag"
  timeout 600 $TR bench.py --gpus 2 --steps 4000 --warmup 20 --no-e2e --no-sweep $flag 2>$O/r2_35.err | python -c "
import sys, json
for l in sys.stdin:
  try: r = json.loads(l)
  except Exception: continue
  print('us/step', r['ms_per_step']*1e3, 'value', r['value'], r.get('shard_check',{}).get('rows_per_rank'))
"
done
tail -2 $O/r2_35.err
echo "== sharded timeline world 2"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 200 python profiles/micro/timeline_sharded.py 32 2 2>&1 | tail -28
