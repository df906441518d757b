#!/bin/bash
# Cluster tail: tree CTA apart from the row CTAs; ncu source view of the kernel.
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_loss_goldens.py -m gpu -x -q > $O/r2_32_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/r2_32_tests.log
timeout 300 python profiles/micro/kernel_times.py --per-graph 10 --batches 32 2>&1 | tail -1
echo "== timeline"; B2R_LIB=profiles/micro/libb200replay_trace.so timeout 100 python profiles/micro/timeline.py 32 2>&1 | tail -24
timeout 100 python profiles/profile_step.py --batch 32 --steps 2 > $O/plain_r2_32.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"c51_post_tree|per_sample_warp" -o $O/prof_r2_chain_b32 -f python profiles/profile_step.py --batch 32 --steps 2 > $O/ncufull_r2_chain_32.log 2>&1
echo "ncu rc=$?"
