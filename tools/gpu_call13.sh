#!/bin/bash
# round-2 call 13: grouped 8x8 convs of ResNeXt through the y-sum kernel (parity + configs[3] leg)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_api.py -m gpu -q -k "resnext or resnet50 or bottleneck" > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2m_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2m_bench.err
CELLSEG_GROUP_YSUM=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2m_bench_nogroup.json 2> /dev/null
python - <<'PY'
import json
for n in ("r2m_bench", "r2m_bench_nogroup"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        x = d["roofline"]["resnext50_32x4d_dense_stride"]
        print(n, "value %.4g frac %.4f clk %s | resnext %.4g inst/s frac %.4f launches %s" % (d["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], x["instances_per_s"], x["frac"], x["launches"]))
    except Exception as e:
        print(n, "unreadable", e)
PY
timeout 300 python profiles/run_arch.py resnext50_32x4d 3 37888 > gpurun_out/r2m_rx_plain.log 2>&1 && cat gpurun_out/r2m_rx_plain.log &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"stem_win|conv_ysum|conv_halo|conv_gemm|head_bf16" --csv --log-file gpurun_out/r2m_rx_launches.csv python profiles/run_arch.py resnext50_32x4d 3 37888 > /dev/null 2>&1
echo "rx ncu rc=$?"
