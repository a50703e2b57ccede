#!/bin/bash
# round-2 call 8 (2 GPUs): select fixes, torchrun MIL-epoch parity script, bench at N = 2 (collectives in the MIL leg)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "select_topk or lexsort or rank" > gpurun_out/r2h_tests_k.log 2>&1; echo "select tests rc=$?"; tail -3 gpurun_out/r2h_tests_k.log
timeout 120 python profiles/time_select.py > gpurun_out/r2h_select_plain.log 2>&1 && cat gpurun_out/r2h_select_plain.log &&
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"select|seg_sort" --csv --log-file gpurun_out/r2h_select_launches.csv python profiles/time_select.py > /dev/null 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_mil_epoch.py > gpurun_out/r2h_dist.log 2>&1; echo "dist rc=$?"
tail -6 gpurun_out/r2h_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 > gpurun_out/r2h_bench2.json 2> gpurun_out/r2h_bench2.err; echo "bench2 rc=$?"
tail -c 600 gpurun_out/r2h_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 --ref-bags 2 > gpurun_out/r2h_ref2.json 2>&1; echo "ref2 rc=$?"
python - <<'PY'
import json
for n in ("r2h_bench2", "r2h_ref2"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % n).read().strip().splitlines() if l.startswith("{")][-1])
        print(n, "value %.4g n_gpus %s" % (d["value"], d["n_gpus"]))
        if "roofline" in d:
            print("   select_20k", {a: b for a, b in d["roofline"]["select_20k"].items() if a not in ("note",)})
        print("   mil", json.dumps(d.get("mil_epoch"))[:1800])
    except Exception as e:
        print(n, "unreadable", e)
PY
