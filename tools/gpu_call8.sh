#!/bin/bash
# round-2 (8 GPUs): torchrun MIL-epoch parity script and the default bench at N = 8
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/dist_mil_epoch.py > gpurun_out/r2q_dist.log 2>&1; echo "dist rc=$?"
grep -E "PASS|FAIL|rank " gpurun_out/r2q_dist.log | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 > gpurun_out/r2q_bench8.json 2> gpurun_out/r2q_bench8.err; echo "bench8 rc=$?"
tail -c 300 gpurun_out/r2q_bench8.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2q_bench8.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("value %.4g n_gpus %s e2e %.4g frac %.4f clk %s" % (d["value"], d["n_gpus"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
    print("   mil", json.dumps(d.get("mil_epoch"))[:1700])
except Exception as e:
    print("unreadable", e)
PY
