#!/bin/bash
# round-2 call 5 (2 GPUs): torchrun MIL-epoch parity script + bench at N = 2 (collectives inside the MIL leg)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_mil_epoch.py > gpurun_out/r2e_dist.log 2>&1; echo "dist rc=$?"
tail -6 gpurun_out/r2e_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err; echo "bench2 rc=$?"
tail -c 600 gpurun_out/r2e_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --workload mil_epoch --steps 2 > gpurun_out/r2e_mil2.json 2> gpurun_out/r2e_mil2.err; echo "mil2 rc=$?"
tail -c 400 gpurun_out/r2e_mil2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 --ref-bags 2 > gpurun_out/r2e_ref2.json 2>&1; echo "ref2 rc=$?"
python - <<'PY'
import json
for n in ("r2e_bench2", "r2e_mil2", "r2e_ref2"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % n).read().strip().splitlines() if l.startswith("{")][-1])
        print(n, "value %.4g n_gpus %s" % (d["value"], d["n_gpus"]))
        print("   mil", json.dumps(d.get("mil_epoch"))[:1500])
    except Exception as e:
        print(n, "unreadable", e)
PY
