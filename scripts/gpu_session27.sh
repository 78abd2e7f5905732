#!/bin/bash
# final captures of the round: tests, the bench line, the ncu launch list and one full capture of the dominant kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s27_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s27_tests.log
tail -3 gpurun_out/s27_tests.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/s27_bench.json 2> gpurun_out/s27_bench.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s27_bench_ref.json 2> gpurun_out/s27_bench_ref.err
echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu --kernels-only > gpurun_out/s27_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_evaluate --launch-skip 3 -c 1 -o gpurun_out/r2f_fused \
  python bench.py --steps 2 --warmup 3 --no-cpu --kernels-only > gpurun_out/s27_ncu.log 2>&1
tail -2 gpurun_out/s27_ncu.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s27_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["frac"], d["clocks"])
print({n: round(v["ms"], 3) for n, v in d["kernels"].items()})
print("latency", {k: (round(v["p50_us"],1) if "p50_us" in v else {kk: round(vv["p50_us"],1) for kk, vv in v.items()}) for k, v in d["replan_latency"].items()})
print("fp32", json.dumps(d["fp32_fast_path"])[:700])
print("literal", json.dumps(d["literal_obstacle_fields"])[:900])
print("default", json.dumps(d["default_lattice_batch"])[:500])
print("closed", json.dumps(d["closed_loop"])[:300])
PY
