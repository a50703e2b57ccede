#!/bin/bash
# round-2 call 6: full tests (dense layer 2 default, select rework), default bench, A/B, launch lists
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2f_tests.log
tail -12 gpurun_out/r2f_tests.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; rc=$?; echo "bench rc=$rc"
tail -c 600 gpurun_out/r2f_bench.err
[ $rc -ne 0 ] && tail -c 800 gpurun_out/r2f_bench.json
Q="--no-side-legs --no-cpu-baseline"
CELLSEG_DENSE_BN=128 timeout 300 python bench.py $Q > gpurun_out/r2f_bn128.json 2>&1
CELLSEG_DENSE_PO=64 timeout 300 python bench.py $Q > gpurun_out/r2f_po64.json 2>&1
timeout 300 python bench.py $Q --max-batch 75776 > gpurun_out/r2f_mb75776.json 2>&1
CELLSEG_YSUM_PAIRS=0 timeout 300 python bench.py $Q > gpurun_out/r2f_nopairs.json 2>&1
python - <<'PY'
import json
for n in ("r2f_bench", "r2f_bn128", "r2f_po64", "r2f_mb75776", "r2f_nopairs"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g e2e %.4g frac %.4f fwd_ms %.2f sel_ms %.4f clk %s verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fwd_ms_per_step"], d["roofline"]["select_in_step"]["ms_per_step"], d["clocks"]["sm_mhz"], (d.get("verify") or {}).get("max_abs_dp_vs_fp32_cuda")))
        if n == "r2f_bench":
            r = d["roofline"]
            for k in ("select_20k", "hsv_refine", "remove_small_regions", "preprocess_masks_chain", "resnext50_32x4d_dense_stride"):
                print("  ", k, {a: b for a, b in r[k].items() if a not in ("workload", "note", "cpu_baseline")})
            print("   mil", json.dumps(d.get("mil_epoch"))[:700])
    except Exception as e:
        print(n, "unreadable", e)
PY
timeout 300 python bench.py --bags-per-step 13 --steps 1 --warmup 1 $Q --no-verify > gpurun_out/r2f_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stem_win|conv_ysum|conv_halo|conv_gemm|head_bf16" -c 140 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --bags-per-step 13 --steps 1 --warmup 1 $Q --no-verify > gpurun_out/r2f_ncu.log 2>&1
echo "ncu rc=$?"
timeout 120 python profiles/time_select.py > gpurun_out/r2f_select_plain.log 2>&1 && cat gpurun_out/r2f_select_plain.log &&
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"select|seg_sort" --csv --log-file gpurun_out/r2f_select_launches.csv python profiles/time_select.py > /dev/null 2>&1
timeout 300 python profiles/run_side_kernels.py > /dev/null 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"unfold" --csv --log-file gpurun_out/r2f_side_launches.csv python profiles/run_side_kernels.py > /dev/null 2>&1
echo "side rc=$?"
