#!/bin/bash
# 8 GPUs: sharded step after the window range finding / one-CTA write-back / deferred copies.
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --gpus 8 --steps 2000 --warmup 20 --no-e2e > $O/r2_24_n8.json 2> $O/r2_24_n8.err; echo "n8 rc=$?"; head -c 1500 $O/r2_24_n8.json; tail -3 $O/r2_24_n8.err
echo; echo "== no wait"
B2R_DEBUG_XCHG_NOWAIT=1 timeout 600 $TR bench.py --gpus 8 --steps 2000 --warmup 20 --no-e2e --no-sweep > $O/r2_24_n8_nowait.json 2>> $O/r2_24_n8.err; echo "rc=$?"; head -c 500 $O/r2_24_n8_nowait.json
