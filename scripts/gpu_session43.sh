#!/bin/bash
# N = 8 only: one collective per sweep (default) vs per step, same box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for mode in sweep perstep; do
  extra=""; [ $mode = perstep ] && extra="--gather-per-step"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu $extra > gpurun_out/s43_n8_$mode.json 2> gpurun_out/s43_n8_$mode.err
  echo "$mode rc=$?"
done
python - <<'PY'
import json
for tag in ("sweep","perstep"):
    d=json.loads(open(f"gpurun_out/s43_n8_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "value %.4g ms %.3f | e2e %.4g ms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["clocks"]["sm_mhz"])
    pr=d["per_rank"]
    print("   ", pr["collective"], pr["sweep_records_consistent"])
    print("    with   ", [round(x,3) for x in pr["ms_per_step_with_gather"]])
    print("    without", [round(x,3) for x in pr["ms_per_step_without_gather"]])
PY
