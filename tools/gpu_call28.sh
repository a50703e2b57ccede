#!/bin/bash
# round-2 call 28: clean-up pass as a programmatic dependent of the select kernel: parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -x -k "(select_topk and (not subprocess or sortplain or staged or persist)) or sample or rank or mil_epoch_single" > gpurun_out/r2ad_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ad_tests.log
AB_ONLY="default,plain launch,ticket" timeout 600 python profiles/time_select_ab.py > gpurun_out/r2ad_select_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2ad_select_ab.log
