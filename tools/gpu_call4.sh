#!/bin/bash
# round-2 call 4: CC two-pass tests, dense-form layer-2 experiment, select / side-kernel profiles
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "remove_small or select_topk_random or hsv or unfold" > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2d_tests.log
Q="--no-side-legs --no-cpu-baseline"
timeout 300 python bench.py $Q > gpurun_out/r2d_base.json 2>&1
CELLSEG_DENSE_PO=16 timeout 300 python bench.py $Q > gpurun_out/r2d_dense16.json 2>&1
timeout 300 python bench.py $Q > gpurun_out/r2d_base2.json 2>&1
CELLSEG_DENSE_PO=16 timeout 300 python bench.py $Q > gpurun_out/r2d_dense16b.json 2>&1
python - <<'PY'
import json
for n in ("r2d_base", "r2d_dense16", "r2d_base2", "r2d_dense16b"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g frac %.4f fwd_ms %.2f clk %s verify %s" % (d["value"], d["roofline"]["frac"], d["roofline"]["fwd_ms_per_step"], d["clocks"]["sm_mhz"], (d.get("verify") or {}).get("max_abs_dp_vs_fp32_cuda")))
    except Exception as e:
        print(n, "unreadable", e, open("gpurun_out/%s.json" % n).read()[-400:])
PY
timeout 120 python profiles/time_select.py > gpurun_out/r2d_select_plain.log 2>&1 && cat gpurun_out/r2d_select_plain.log &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"select|seg_sort" --csv --log-file gpurun_out/r2d_select_launches.csv python profiles/time_select.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_reg" -s 2 -c 1 -o gpurun_out/r2d_select_reg python profiles/time_select.py > gpurun_out/r2d_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 300 python profiles/run_side_kernels.py > gpurun_out/r2d_side_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"unfold|remove_small|paint|hsv_refine|heat|rank" --csv --log-file gpurun_out/r2d_side_launches.csv python profiles/run_side_kernels.py > /dev/null 2>&1
echo "side rc=$?"
