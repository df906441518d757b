#!/bin/bash
# Early publish of shard totals: emulated-rank parity tests, then 2 GPUs with and without.
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_parity.py -m gpu -x -q -k "shard" > $O/r2_34_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/r2_34_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
for flag in "" "--no-early-publish"; do
  echo "== N=2 $flag"
  timeout 600 $TR bench.py --gpus 2 --steps 4000 --warmup 20 --no-e2e --no-sweep $flag 2>$O/r2_34.err | python -c "
import sys, json
for l in sys.stdin:
  try: r = json.loads(l)
  except Exception: continue
  print('us/step', r['ms_per_step']*1e3, 'value', r['value'], r.get('shard_check',{}).get('rows_per_rank'))
"
done
tail -2 $O/r2_34.err
echo "== N=1"; timeout 300 python bench.py --steps 4000 --warmup 20 --no-e2e --no-cpu-baseline --no-sweep | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('us/step', r['ms_per_step']*1e3)"
