#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -4
bash profiles/r2/run49.sh
