#!/bin/bash
# C51 loss in two halves (pre beside the sampler, tail + write-back in one CTA): suite, times.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_15_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/r2_15_tests.log
timeout 300 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_15_kt.log 2>&1; cat $O/r2_15_kt.log
echo "== unsplit"; B2R_C51_SPLIT=0 timeout 300 python profiles/micro/kernel_times.py --per-graph 10 2>&1 | tail -4
timeout 300 python profiles/micro/c51_error.py --rows 4096 --seeds 2 > $O/r2_15_c51err.log 2>&1; cat $O/r2_15_c51err.log
