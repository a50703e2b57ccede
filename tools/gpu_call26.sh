#!/bin/bash
# round-2 call 26: select kernel in two-warp CTAs: parity (default + variants), A/B timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "select_topk and (not subprocess or cta64 or occ12)" > gpurun_out/r2ab_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ab_tests.log
timeout 600 python profiles/time_select_ab.py > gpurun_out/r2ab_select_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2ab_select_ab.log
