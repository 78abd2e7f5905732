#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s44_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s44_tests.log
tail -3 gpurun_out/s44_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/s44_bench.json 2> gpurun_out/s44_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s44_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f e2e %.4g (%.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["roofline"]["frac"], d["clocks"])
print({n: round(v["ms"], 3) for n, v in d["kernels"].items()})
print("latency", {k: (round(v["p50_us"],1) if "p50_us" in v else {kk: round(vv["p50_us"],1) for kk, vv in v.items()}) for k, v in d["replan_latency"].items()})
print("spot", {k: (v["n"], v["mismatches"]) for k, v in d["parity_spot_check"].items()})
PY
