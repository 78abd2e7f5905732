#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "pipelined" > gpurun_out/s32_tests.log 2>&1; tail -3 gpurun_out/s32_tests.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --kernels-only > gpurun_out/s32_n1.json 2> gpurun_out/s32_n1.err
for mode in sweep perstep; do
  extra=""; [ $mode = perstep ] && extra="--gather-per-step"
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu $extra > gpurun_out/s32_n2_$mode.json 2> gpurun_out/s32_n2_$mode.err
  echo "$mode rc=$?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/s32_n2_$mode.err | tail -3
done
python - <<'PY'
import json
v1=None
for tag in ("n1","n2_sweep","n2_perstep"):
    try:
        d=json.loads(open(f"gpurun_out/s32_{tag}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(tag, "failed", e); continue
    if tag=="n1": v1=d["value"]
    n=d["n_gpus"]
    print(tag, "value %.4g ms %.3f e2e %.4g eff %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["value"]/(n*v1)), d["clocks"]["sm_mhz"])
    if d.get("per_rank"):
        for k, v in d["per_rank"].items():
            print("   ", k, [round(x, 3) for x in v] if isinstance(v, list) else v)
PY
