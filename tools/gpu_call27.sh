#!/bin/bash
# round-2 call 27: ncu --set full with source of the default select kernel (20 000 bags x 3025)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_reg_kernel -s 3 -c 1 -o gpurun_out/prof_select_r02 python profiles/time_select.py > gpurun_out/r2ac_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2ac_ncu.log; ls -la gpurun_out/prof_select_r02.ncu-rep
