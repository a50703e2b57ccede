#!/bin/bash
# round-2 call 1: full GPU tests, default bench, A/B of the y-sum epilogue, launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -15 gpurun_out/r2a_tests.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; rc=$?; echo "bench rc=$rc"
tail -c 600 gpurun_out/r2a_bench.err
if [ $rc -ne 0 ]; then
  tail -c 600 gpurun_out/r2a_bench.json
  CELLSEG_SELECT_FAST=staged timeout 600 python bench.py > gpurun_out/r2a_bench_staged.json 2> gpurun_out/r2a_bench_staged.err; echo "bench(staged select) rc=$?"
  tail -c 1500 gpurun_out/r2a_bench_staged.json
fi
CELLSEG_YSUM_EPI=8 timeout 300 python bench.py --no-side-legs --no-cpu-baseline --no-verify > gpurun_out/r2a_bench_epi8.json 2>&1
CELLSEG_YSUM_PAIRS=1 timeout 300 python bench.py --no-side-legs --no-cpu-baseline --no-verify > gpurun_out/r2a_bench_pairs.json 2>&1
python - <<'PY'
import json
for n in ("r2a_bench", "r2a_bench_epi8", "r2a_bench_pairs"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.4g e2e %.4g frac %.4f fwd_ms %.2f verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fwd_ms_per_step"], d.get("verify")))
    except Exception as e:
        print(n, "unreadable", e)
PY
timeout 300 python bench.py --bags-per-step 13 --steps 1 --warmup 1 --no-side-legs --no-cpu-baseline --no-verify > gpurun_out/r2a_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --bags-per-step 13 --steps 1 --warmup 1 --no-side-legs --no-cpu-baseline --no-verify > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?"
