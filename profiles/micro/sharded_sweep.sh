#!/bin/bash
# Sharded step (peer exchange) at several per-GPU batch sizes; usage: sharded_sweep.sh N
N=${1:-2}
for b in 32 128 512 2048; do
  timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port $((29600 + b % 97)) bench.py --gpus $N \
    --batch $b --steps 500 --warmup 20 --no-e2e 2>/dev/null | grep '^{' | \
    B=$b python -c "
import json, os, sys
l = json.loads(sys.stdin.read().splitlines()[-1])
print('batch/gpu', os.environ['B'], 'global', l['config']['workload'].split('global batch ')[-1],
      'transitions/s', l['value'], 'ms/step', l['ms_per_step'])"
done
