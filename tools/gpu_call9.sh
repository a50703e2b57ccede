#!/bin/bash
# round-2 call 9: persistent select + single-chunk offsets scan
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "select_topk or lexsort or rank" > gpurun_out/r2i_tests_k.log 2>&1; echo "select tests rc=$?"; tail -3 gpurun_out/r2i_tests_k.log
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -q -k "mil_epoch or sample or train_tile" > gpurun_out/r2i_tests_a.log 2>&1; echo "api tests rc=$?"; tail -3 gpurun_out/r2i_tests_a.log
for v in 1 0; do
  CELLSEG_SELECT_PERSIST=$v timeout 120 python profiles/time_select.py > gpurun_out/r2i_select_plain_$v.log 2>&1 && cat gpurun_out/r2i_select_plain_$v.log &&
  CELLSEG_SELECT_PERSIST=$v timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"select|seg_sort" --csv --log-file gpurun_out/r2i_select_launches_$v.csv python profiles/time_select.py > /dev/null 2>&1
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; rc=$?; echo "bench rc=$rc"
tail -c 400 gpurun_out/r2i_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g frac %.4f sel_ms %.4f clk %s verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["select_in_step"]["ms_per_step"], d["clocks"]["sm_mhz"], (d.get("verify") or {}).get("selection_equals_oracle")))
    print("  select_20k", {a: b for a, b in d["roofline"]["select_20k"].items() if a != "note"})
    print("  mil", json.dumps(d.get("mil_epoch"))[:600])
except Exception as e:
    print("unreadable", e)
PY
