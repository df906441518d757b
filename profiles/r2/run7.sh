#!/bin/bash
set -u
O=gpurun_out
timeout 200 python profiles/micro/kernel_times.py --per-graph 10 > $O/r2_7_kt.log 2>&1; cat $O/r2_7_kt.log
B2R_GATHER_PAD_KB=0 timeout 200 python profiles/micro/kernel_times.py --per-graph 10 --batches 256,1024 2>&1 | tail -2
