#!/bin/bash
# round-2 call 37: the select variants not re-run since the clean-up pass became a programmatic dependent
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -n 3 -k "other_paths and (cta64 or occ16 or ticket)" > gpurun_out/r2al_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2al_tests.log
