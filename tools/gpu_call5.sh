#!/bin/bash
# round-2 (2 GPUs): torchrun MIL-epoch parity script, bench at N = 2 (collectives in the MIL leg), reference arm under torchrun
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_mil_epoch.py > gpurun_out/r2p_dist.log 2>&1; echo "dist rc=$?"
grep -E "PASS|FAIL|rank " gpurun_out/r2p_dist.log | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 > gpurun_out/r2p_bench2.json 2> gpurun_out/r2p_bench2.err; echo "bench2 rc=$?"
tail -c 300 gpurun_out/r2p_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --workload mil_epoch --steps 2 > gpurun_out/r2p_mil2.json 2> gpurun_out/r2p_mil2.err; echo "mil2 rc=$?"
python - <<'PY'
import json
for n in ("r2p_bench2", "r2p_mil2"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % n).read().strip().splitlines() if l.startswith("{")][-1])
        print(n, "value %.4g n_gpus %s e2e %.4g" % (d["value"], d["n_gpus"], d["e2e"]["value"]))
        print("   mil", json.dumps((d.get("mil_epoch") or {}).get("cache_features"))[:900])
    except Exception as e:
        print(n, "unreadable", e)
PY
