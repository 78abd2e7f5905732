#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== new tests" > gpurun_out/s7_tests.log
timeout 900 python -m pytest tests -m gpu -x -q -k "fp32 or optimal_targets or follow or error_codes or odd_sizes" >> gpurun_out/s7_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s7_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err
echo "bench rc=$?" >> gpurun_out/s7_tests.log
timeout 300 python scripts/latency_breakdown.py > gpurun_out/s7_latency_breakdown.txt 2>&1
timeout 300 python scripts/profile_replan_python.py > gpurun_out/s7_profile_python.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --kernels-only > gpurun_out/s7_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_evaluate --launch-skip 3 -c 1 -o gpurun_out/r2_fused_table \
  python bench.py --steps 2 --warmup 3 --no-cpu --kernels-only > gpurun_out/s7_ncu.log 2>&1
tail -4 gpurun_out/s7_tests.log
cat gpurun_out/s7_latency_breakdown.txt
head -30 gpurun_out/s7_profile_python.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s7_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), d["roofline"]["frac"], {n: round(v["ms"], 3) for n, v in d["kernels"].items()})
print("fp32", json.dumps(d["fp32_fast_path"])[:900])
print("latency", {k: v["p50_us"] for k, v in d["replan_latency"].items()})
PY
