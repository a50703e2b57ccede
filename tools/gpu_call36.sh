#!/bin/bash
# round-2 call 36 (tabulated ranges, bag table in shared memory): heat-map painting as a gather: parity (kernel + heatmap() flow), timing against the scatter form
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -m gpu -q -x -k "paint or heatmap" > gpurun_out/r2ak_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2ak_tests.log
timeout 300 python profiles/time_paint.py > gpurun_out/r2ak_paint.log 2>&1; echo "paint rc=$?"; cat gpurun_out/r2ak_paint.log
