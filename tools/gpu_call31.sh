#!/bin/bash
# round-2 call 31: 32-column threshold for small kept counts: parity + A/B; >31 744-bag test
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -x -k "(select_topk and (not subprocess or col64 or recount or cta64occ16)) or sample or rank or mil_epoch_single" > gpurun_out/r2ag_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ag_tests.log
AB_ONLY="default,64 columns" timeout 600 python profiles/time_select_ab.py > gpurun_out/r2ag_select_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2ag_select_ab.log
