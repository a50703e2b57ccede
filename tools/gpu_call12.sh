#!/bin/bash
# round-2 call 12: specialised head kernel (parity + bench), ResNeXt-50 per-launch profile
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_api.py -m gpu -q -k "not subprocess and not many_iterations and not conv_matches" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2l_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2l_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g frac %.4f fwd_ms %.2f clk %s verify %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["fwd_ms_per_step"], d["clocks"]["sm_mhz"], (d.get("verify") or {}).get("max_abs_dp_vs_fp32_cuda")))
except Exception as e:
    print("unreadable", e)
PY
timeout 300 python profiles/run_arch.py resnext50_32x4d 3 37888 > gpurun_out/r2l_rx_plain.log 2>&1 && cat gpurun_out/r2l_rx_plain.log &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"stem_win|conv_ysum|conv_halo|conv_gemm|head_bf16" --csv --log-file gpurun_out/r2l_rx_launches.csv python profiles/run_arch.py resnext50_32x4d 3 37888 > /dev/null 2>&1
echo "rx ncu rc=$?"
