#!/bin/bash
# Early-publish tests (emulated ranks) and the complete default bench on 8 GPUs, both arms.
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_abi_and_host.py -m gpu -x -q -k "early or shard or abi" > $O/r2_37_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/r2_37_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_37_n8.json 2> $O/r2_37_n8.err; echo "n8 rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_37_n8.json'):
  if l.startswith('{'):
    r = json.loads(l)
    for k in ('value', 'ms_per_step', 'sweep_summary', 'e2e', 'shard_check', 'timing'):
      print(k, str(r.get(k))[:500])
PY
tail -3 $O/r2_37_n8.err
timeout 600 $TR bench.py --impl reference --gpus 8 --steps 5 --warmup 3 > $O/r2_37_ref_n8.json 2> $O/r2_37_ref.err; echo "ref rc=$?"; head -c 300 $O/r2_37_ref_n8.json
