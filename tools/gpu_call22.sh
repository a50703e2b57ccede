#!/bin/bash
# round-2 call 22: four-set epilogue ring for the residual GEMMs: parity, then A/B against the previous build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=cellsegmentation_b200/csrc/libcellseg_b200.so
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "(basic_block or within_2e2 or bench_scale or conv_matches or tile16) and not subprocess" > gpurun_out/r2y_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -5 gpurun_out/r2y_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
cp $L /tmp/lib_new.so
for round in 1 2; do
  for v in prev new; do
    if [ $v = prev ]; then cp tools/ab/lib_prev.so $L; else cp /tmp/lib_new.so $L; fi
    timeout 400 python bench.py --no-side-legs --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/r2y_bench_$v$round.json 2> gpurun_out/r2y_bench_$v$round.err; echo "$v rc=$?"
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_bench_$v$round.json").read().strip().splitlines()[-1])
print("$v", "value %.4g e2e %.4g frac %.4f clk %s verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], (d.get("verify") or {}).get("max_abs_dp_vs_fp32_cuda")))
PY
  done
done
cp /tmp/lib_new.so $L
