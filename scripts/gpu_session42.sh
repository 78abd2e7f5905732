#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "pipelined" > gpurun_out/s42_tests.log 2>&1; tail -3 gpurun_out/s42_tests.log
for rep in 1 2; do
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --kernels-only > gpurun_out/s42_n1_$rep.json 2> gpurun_out/s42_n1_$rep.err
done
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/s42_n2.json 2> gpurun_out/s42_n2.err
echo "n2 rc=$?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/s42_n2.err | tail -3
python - <<'PY'
import json
for tag in ("n1_1","n1_2","n2"):
    d=json.loads(open(f"gpurun_out/s42_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "value %.4g ms %.3f | e2e %.4g ms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["survivors"])
PY
