#!/bin/bash
# Session re-entry baseline: GPU suite, default bench line, per-kernel times.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_14_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_14_tests.log
timeout 600 python bench.py > $O/r2_14_bench.json 2> $O/r2_14_bench.err; echo "bench rc=$?"; head -c 3000 $O/r2_14_bench.json
timeout 300 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_14_kt.log 2>&1; cat $O/r2_14_kt.log
