#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_other_learners.py tests/test_learner.py -m gpu -x -q > $O/r2_42_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/r2_42_tests.log
