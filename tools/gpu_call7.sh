#!/bin/bash
# round-2 call 7: select v3 + offsets prefetch, new default forward batch (75 776)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q > gpurun_out/r2g_tests_k.log 2>&1; echo "kernel tests rc=$?"; tail -4 gpurun_out/r2g_tests_k.log
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_api.py -m gpu -q -k "bench_scale or mil_epoch or sample or rank" > gpurun_out/r2g_tests_m.log 2>&1; echo "model tests rc=$?"; tail -4 gpurun_out/r2g_tests_m.log
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; rc=$?; echo "bench rc=$rc"
tail -c 600 gpurun_out/r2g_bench.err
[ $rc -ne 0 ] && tail -c 800 gpurun_out/r2g_bench.json
python - <<'PY'
import json
for n in ("r2g_bench",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g e2e %.4g frac %.4f fwd_ms %.2f sel_ms %.4f clk %s verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fwd_ms_per_step"], d["roofline"]["select_in_step"]["ms_per_step"], d["clocks"]["sm_mhz"], (d.get("verify") or {}).get("max_abs_dp_vs_fp32_cuda")))
        r = d["roofline"]
        for k in ("select_20k", "remove_small_regions", "preprocess_masks_chain"):
            print("  ", k, {a: b for a, b in r[k].items() if a not in ("workload", "note", "cpu_baseline")})
    except Exception as e:
        print(n, "unreadable", e)
PY
timeout 120 python profiles/time_select.py > gpurun_out/r2g_select_plain.log 2>&1 && cat gpurun_out/r2g_select_plain.log &&
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"select|seg_sort" --csv --log-file gpurun_out/r2g_select_launches.csv python profiles/time_select.py > /dev/null 2>&1
echo "select ncu rc=$?"
