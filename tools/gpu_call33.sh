#!/bin/bash
# round-2 call 33: select paths after the clean-up-pass ordering fix (persistent kernel waits on every path; PDL clean-up only behind kernels that do)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -x -k "(select_topk and (not subprocess or persist or staged or warp)) or sample or rank or mil_epoch_single" > gpurun_out/r2ah_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ah_tests.log
