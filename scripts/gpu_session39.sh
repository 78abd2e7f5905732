#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/s39_bench.json 2> gpurun_out/s39_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s39_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["frac"], d["clocks"])
print({n: round(v["ms"], 3) for n, v in d["kernels"].items()})
print("latency", {k: (round(v["p50_us"],1) if "p50_us" in v else {kk: round(vv["p50_us"],1) for kk, vv in v.items()}) for k, v in d["replan_latency"].items()})
print("spot", {k: (v["n"], v["mismatches"]) for k, v in d["parity_spot_check"].items()})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
PY
