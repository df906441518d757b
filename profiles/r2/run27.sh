#!/bin/bash
# Host side of one end-to-end update: eager launches against the graph-replayed step.
set -u
O=gpurun_out
echo "== eager"; B2R_HOST_TRACE=1 timeout 300 python profiles/micro/e2e_breakdown.py 2>&1 | tail -60
echo "== graph"; B2R_TRAINER_GRAPH=1 B2R_HOST_TRACE=1 timeout 300 python profiles/micro/e2e_breakdown.py 2>&1 | tail -60
timeout 300 python profiles/micro/host_batch_breakdown.py 32 2>&1 | tail -12
