#!/bin/bash
# round-2 final 1-GPU call: full GPU test suite, smoke, default bench + reference arm, launch list and
# `ncu --set full` capture of one forward batch (profiles/r02_*), select timing, forward-batch A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2z_tests.log
tail -6 gpurun_out/r2z_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2z_bench.err
Q="--no-side-legs --no-cpu-baseline --no-verify --bags-per-step 26 --steps 1 --warmup 1"
K='regex:stem_ts|stem_win|ysum_block|conv_ysum|conv_halo|conv_gemm|head_bf16'
timeout 300 python bench.py $Q > gpurun_out/r2z_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 130 --csv --log-file gpurun_out/r2z_launches.csv python bench.py $Q > gpurun_out/r2z_ncu.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 62 -c 31 -o gpurun_out/prof_r02 python bench.py $Q > gpurun_out/r2z_ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_r02.ncu-rep
timeout 300 python bench.py --no-side-legs --no-cpu-baseline --max-batch 151552 > gpurun_out/r2z_bench_mb151552.json 2> gpurun_out/r2z_bench_mb151552.err; echo "mb151552 rc=$?"
timeout 300 python bench.py --no-side-legs --no-cpu-baseline > gpurun_out/r2z_bench_mb75776.json 2> gpurun_out/r2z_bench_mb75776.err; echo "mb75776 rc=$?"
timeout 400 python profiles/time_select_ab.py > gpurun_out/r2z_select_ab.log 2>&1; echo "select ab rc=$?"; cat gpurun_out/r2z_select_ab.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2>&1; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("r2z_bench", "r2z_bench_mb151552", "r2z_bench_mb75776"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "value %.4g e2e %.4g frac %.4f clk %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]),
              "select_20k", (d["roofline"].get("select_20k") or {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
