#!/bin/bash
# round-2 call 14: grouped dense-form experiment for the 4x4 stage of ResNeXt
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in 4 16 4 16; do
  echo "CELLSEG_DENSE_GROUP_PO=$v"; CELLSEG_DENSE_GROUP_PO=$v timeout 300 python profiles/run_arch.py resnext50_32x4d 3 75776 2>&1 | tail -1
done
CELLSEG_DENSE_GROUP_PO=16 timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -k "(resnext or conv_matches) and not subprocess" > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2n_tests.log
