#!/bin/bash
# 2 GPUs: sharded step sized by the local share; what the cross-GPU coupling costs.
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 2000 --warmup 20 --no-e2e > $O/r2_21_n2.json 2> $O/r2_21_n2.err; echo "n2 rc=$?"; head -c 1800 $O/r2_21_n2.json; tail -3 $O/r2_21_n2.err
echo; echo "== no wait"
B2R_DEBUG_XCHG_NOWAIT=1 timeout 600 $TR bench.py --gpus 2 --steps 2000 --warmup 20 --no-e2e > $O/r2_21_n2_nowait.json 2>> $O/r2_21_n2.err; echo "rc=$?"; head -c 700 $O/r2_21_n2_nowait.json
echo; echo "== N=1 same flags"
timeout 300 python bench.py --steps 2000 --warmup 20 --no-e2e --no-cpu-baseline --no-sweep | head -c 400
