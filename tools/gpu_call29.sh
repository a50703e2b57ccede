#!/bin/bash
# round-2 call 29: recount offsets kernel (no memset, no inter-CTA traffic): parity + A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -x -k "(select_topk and (not subprocess or lookback or ticket or warp)) or sample or rank or mil_epoch_single" > gpurun_out/r2ae_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ae_tests.log
AB_ONLY="default,lookback offsets,plain launch,ticket" timeout 600 python profiles/time_select_ab.py > gpurun_out/r2ae_select_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2ae_select_ab.log
