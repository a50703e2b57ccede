#!/bin/bash
# round-2 call 25: default bench line of the final tree (select leg timed hot and after idle)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2zz_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2zz_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.4g e2e %.4g frac %.4f clk %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
print(json.dumps(d["roofline"]["select_20k"]))
print(d["roofline"]["traffic_detail"])
PY
