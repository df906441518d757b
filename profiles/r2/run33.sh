#!/bin/bash
# Full suite + default bench.
set -u
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2_33_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/r2_33_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > $O/r2_33_bench.json 2> $O/r2_33_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
r = json.load(open('gpurun_out/r2_33_bench.json'))
for k in ('value', 'ms_per_step', 'sweep_summary', 'e2e', 'e2e_host_batch', 'cpu_baseline'):
  print(k, str(r.get(k))[:400])
PY
tail -3 $O/r2_33_bench.err
