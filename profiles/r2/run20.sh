#!/bin/bash
# Full suite + default bench line after: loss in two halves, deferred copies, max_rows.
set -u
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_20_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/r2_20_tests.log
timeout 900 python bench.py > $O/r2_20_bench.json 2> $O/r2_20_bench.err; echo "bench rc=$?"; head -c 2500 $O/r2_20_bench.json; tail -3 $O/r2_20_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2_20_bench_k20.json 2>> $O/r2_20_bench.err; echo; head -c 1500 $O/r2_20_bench_k20.json
