#!/bin/bash
# round-2 call 17: fused layer-1 BasicBlock kernel: parity alone, forward tests, A/B against two y-sum launches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "basic_block" > gpurun_out/r2t_block.log 2>&1; rc=$?; echo "block rc=$rc"; tail -25 gpurun_out/r2t_block.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -k "(within_2e2 or bench_scale or tile16) and not subprocess" > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2t_tests.log
for round in 1 2; do
  for v in 0 1; do
    CELLSEG_BLOCK_FUSE=$v timeout 400 python bench.py --no-side-legs --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/r2t_bench_$v$round.json 2> gpurun_out/r2t_bench_$v$round.err; echo "fuse=$v rc=$?"
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2t_bench_$v$round.json").read().strip().splitlines()[-1])
print("fuse=$v", "value %.4g e2e %.4g frac %.4f clk %s launches %s verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d.get("gpu_launches"), (d.get("verify") or {}).get("max_abs_dp_vs_fp32_cuda")))
PY
  done
done
