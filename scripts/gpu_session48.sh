#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "pipelined" > gpurun_out/s48_tests.log 2>&1; tail -2 gpurun_out/s48_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/s48_n2.json 2> gpurun_out/s48_n2.err
echo "n2 rc=$?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/s48_n2.err | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s48_n2.json").read().strip().splitlines()[-1])
print("value %.4g ms %.3f | e2e %.4g ms %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["survivors"], d["per_rank"]["collective"], d["per_rank"]["sweep_records_consistent"])
PY
