#!/bin/bash
# round-2 call 11: one-box y-sum (probe, parity, A/B) + head loop
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 ./tools/umma_probe > gpurun_out/r2k_probe.log 2>&1; echo "probe rc=$?"; grep -A0 "^==\|match" gpurun_out/r2k_probe.log | tail -16
CELLSEG_YSUM_BOX=1 timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -k "(conv_matches or many_iterations or within_2e2 or bench_scale) and not subprocess" > gpurun_out/r2k_tests.log 2>&1; echo "one-box tests rc=$?"; tail -4 gpurun_out/r2k_tests.log
Q="--no-side-legs --no-cpu-baseline"
for v in 1 0 1 0; do
  CELLSEG_YSUM_BOX=$v timeout 300 python bench.py $Q > gpurun_out/r2k_box${v}_$RANDOM.json 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2k_box*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g e2e %.4g frac %.4f fwd_ms %.2f clk %s verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fwd_ms_per_step"], d["clocks"]["sm_mhz"], (d.get("verify") or {}).get("max_abs_dp_vs_fp32_cuda")))
    except Exception as e:
        print(f, "unreadable", e, open(f).read()[-300:])
PY
