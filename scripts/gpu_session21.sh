#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s21_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s21_tests.log
tail -4 gpurun_out/s21_tests.log
timeout 1500 python bench.py > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s21_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s21_bench_ref.json 2> gpurun_out/s21_bench_ref.err
echo "ref rc=$?"; tail -c 600 gpurun_out/s21_bench_ref.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s21_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["frac"], {n: round(v["ms"], 3) for n, v in d["kernels"].items()})
print("latency", json.dumps(d["replan_latency"])[:1500])
print("spot", d["parity_spot_check"])
print("cpu", json.dumps(d["cpu_baseline"])[:600])
print("default_lattice_batch", json.dumps(d["default_lattice_batch"])[:700])
PY
