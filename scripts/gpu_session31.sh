#!/bin/bash
# the scaling run the driver does at round end (N = 1, 2, 4, 8 on one 8-GPU box), plus N = 8 with speed-balanced shards
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --kernels-only > gpurun_out/scale2_n1.json 2> gpurun_out/scale2_n1.err
for n in 2 4 8; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu > gpurun_out/scale2_n$n.json 2> gpurun_out/scale2_n$n.err
  echo "n$n rc=$?"
done
true

python - <<'PY'
import json
v1=None
for tag in ("n1","n2","n4","n8"):
    try:
        d=json.loads(open(f"gpurun_out/scale2_{tag}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(tag, "failed", e); continue
    if tag=="n1": v1=d["value"]
    n=d["n_gpus"]
    print(tag, "value %.4g ms %.3f e2e %.4g eff %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["value"]/(n*v1)), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    if d.get("per_rank"):
        for k, v in d["per_rank"].items():
            if isinstance(v, list): print("   ", k, [round(x, 3) for x in v])
PY
