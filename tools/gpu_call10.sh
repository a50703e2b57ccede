#!/bin/bash
# round-2 call 10: stem + layer 1 in L2-resident sub-batches (A/B over the sub-batch size), two-deep e2e pipeline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -k "bench_scale or within_2e2 or tile16 or hilo" > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2j_tests.log
Q="--no-side-legs --no-cpu-baseline"
for sub in 4736 0 9472 2368 4736 0; do
  CELLSEG_L1_SUB=$sub timeout 300 python bench.py $Q > gpurun_out/r2j_sub${sub}_$RANDOM.json 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2j_sub*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g frac %.4f fwd_ms %.2f clk %s verify %s launches %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fwd_ms_per_step"], d["clocks"]["sm_mhz"], (d.get("verify") or {}).get("max_abs_dp_vs_fp32_cuda"), d["gpu_launches"]))
    except Exception as e:
        print(f, "unreadable", e, open(f).read()[-300:])
PY
