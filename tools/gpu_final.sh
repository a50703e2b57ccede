#!/bin/bash
# round-2 final 1-GPU call: full GPU test suite, smoke, default bench + reference arm, launch list and
# `ncu --set full` capture of one forward batch (profiles/r02_*)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2z_tests.log
tail -6 gpurun_out/r2z_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2>&1; echo "ref rc=$?"
Q="--no-side-legs --no-cpu-baseline --no-verify --bags-per-step 26 --steps 1 --warmup 1"
timeout 300 python bench.py $Q > gpurun_out/r2z_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stem_ts|stem_win|ysum_block|conv_ysum|conv_halo|conv_gemm|head_bf16" -c 130 --csv --log-file gpurun_out/r2z_launches.csv python bench.py $Q > gpurun_out/r2z_ncu.log 2>&1
echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none -k regex:"stem_ts|stem_win|ysum_block|conv_ysum|conv_halo|conv_gemm|head_bf16" -s 62 -c 31 -o gpurun_out/prof_r02 python bench.py $Q > gpurun_out/r2z_ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_r02.ncu-rep
