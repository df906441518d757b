#!/bin/bash
# Suite + bench after: add_batch, network-input kernel, TF Adam, gather variant by batch size.
set -u
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2_29_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2_29_tests.log
timeout 900 python bench.py --no-cpu-baseline > $O/r2_29_bench.json 2> $O/r2_29_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
r = json.load(open('gpurun_out/r2_29_bench.json'))
for k in ('value', 'ms_per_step', 'sweep_summary', 'e2e', 'e2e_sync', 'e2e_host_batch', 'full_train_step', 'full_train_step_graph'):
  print(k, r.get(k))
PY
tail -3 $O/r2_29_bench.err
