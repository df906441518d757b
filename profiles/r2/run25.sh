#!/bin/bash
# Suite + bench after: pinned host batches, trainer loss double-buffering, tree geometry by expected rows.
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_25_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2_25_tests.log
timeout 900 python bench.py --no-cpu-baseline > $O/r2_25_bench.json 2> $O/r2_25_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
r = json.load(open('gpurun_out/r2_25_bench.json'))
for k in ('value', 'ms_per_step', 'sweep_summary', 'e2e', 'e2e_sync', 'e2e_host_batch'):
  print(k, r.get(k))
PY
tail -3 $O/r2_25_bench.err
