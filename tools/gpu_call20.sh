#!/bin/bash
# round-2 call 20: full ncu capture of one forward batch of the current build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
Q="--no-side-legs --no-cpu-baseline --no-verify --bags-per-step 26 --steps 1 --warmup 1"
K='regex:stem_ts|stem_win|ysum_block|conv_ysum|conv_halo|conv_gemm|head_bf16'
timeout 300 python bench.py $Q > gpurun_out/r2w_plain.log 2>&1 &&
timeout 1500 ncu --set full --import-source on --clock-control none -k "$K" -s 62 -c 31 -o gpurun_out/prof_r02w python bench.py $Q > gpurun_out/r2w_ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_r02w.ncu-rep
