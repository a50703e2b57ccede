#!/bin/bash
# round-2 call 3: full GPU tests, default bench, A/B of y-sum pairs and the fused layer-2 shortcut, launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2c_tests.log
tail -25 gpurun_out/r2c_tests.log
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; rc=$?; echo "bench rc=$rc"
tail -c 800 gpurun_out/r2c_bench.err
[ $rc -ne 0 ] && tail -c 800 gpurun_out/r2c_bench.json
Q="--no-side-legs --no-cpu-baseline"
CELLSEG_YSUM_PAIRS=0 timeout 300 python bench.py $Q > gpurun_out/r2c_nopairs.json 2>&1
CELLSEG_HALO_DS=0 timeout 300 python bench.py $Q > gpurun_out/r2c_nohalods.json 2>&1
timeout 300 python bench.py $Q --max-batch 75776 > gpurun_out/r2c_mb75776.json 2>&1
python - <<'PY'
import json
for n in ("r2c_bench", "r2c_nopairs", "r2c_nohalods", "r2c_mb75776"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g e2e %.4g frac %.4f fwd_ms %.2f sel_ms %.4f clk %s verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fwd_ms_per_step"], d["roofline"]["select_in_step"]["ms_per_step"], d["clocks"]["sm_mhz"], d.get("verify")))
        if n == "r2c_bench":
            r = d["roofline"]
            for k in ("select_20k", "hsv_refine", "remove_small_regions", "preprocess_masks_chain", "resnext50_32x4d_dense_stride"):
                print("  ", k, {a: b for a, b in r[k].items() if a not in ("workload", "note", "cpu_baseline")}, r[k].get("cpu_baseline", {}).get("value"))
            print("   mil", json.dumps(d.get("mil_epoch"))[:1800])
            print("   cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(n, "unreadable", e)
PY
timeout 300 python bench.py --bags-per-step 13 --steps 1 --warmup 1 $Q --no-verify > gpurun_out/r2c_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stem_win|conv_|head_bf16|select|seg_sort" -c 300 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --bags-per-step 13 --steps 1 --warmup 1 $Q --no-verify > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?"
