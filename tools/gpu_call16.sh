#!/bin/bash
# round-2 call 16: weights-stationary stem: parity alone, then the forward tests, then A/B against the window form
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "stem_matches" > gpurun_out/r2s_stem.log 2>&1; rc=$?; echo "stem rc=$rc"; tail -15 gpurun_out/r2s_stem.log
CELLSEG_STEM=win timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "stem_matches" > gpurun_out/r2s_stem_win.log 2>&1; echo "stem(win) rc=$?"; tail -3 gpurun_out/r2s_stem_win.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -k "(within_2e2 or bench_scale or tile16) and not subprocess" > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2s_tests.log
for round in 1 2; do
  for v in win ts; do
    CELLSEG_STEM=$v timeout 400 python bench.py --no-side-legs --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/r2s_bench_$v$round.json 2> gpurun_out/r2s_bench_$v$round.err; echo "$v rc=$?"
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2s_bench_$v$round.json").read().strip().splitlines()[-1])
print("$v", "value %.4g e2e %.4g frac %.4f clk %s verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], (d.get("verify") or {}).get("max_abs_dp_vs_fp32_cuda")))
PY
  done
done
