#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --kernels-only > gpurun_out/s33_n1.json 2> gpurun_out/s33_n1.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s33_n2.json 2> gpurun_out/s33_n2.err
echo "n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/s33_ref_n2.json 2> gpurun_out/s33_ref_n2.err
echo "ref n2 rc=$?"; tail -c 400 gpurun_out/s33_ref_n2.json
python - <<'PY'
import json
for tag in ("n1","n2"):
    d=json.loads(open(f"gpurun_out/s33_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), d["clocks"])
PY
