#!/bin/bash
# round-2 call 23: full GPU suite on the build with the warp-per-bag select + look-back offsets,
# A/B timing of the select variants, ncu metric pass of the selection launches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2aa_tests.log
tail -25 gpurun_out/r2aa_tests.log
timeout 600 python profiles/time_select_ab.py > gpurun_out/r2aa_select_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2aa_select_ab.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"select_|seg_sort" --csv --log-file gpurun_out/r2aa_select_ncu.csv python profiles/time_select.py > gpurun_out/r2aa_select_ncu.log 2>&1; echo "ncu rc=$?"
tail -12 gpurun_out/r2aa_select_ncu.csv | cut -c1-400
