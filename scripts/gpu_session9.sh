#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "pipelined or fp32" > gpurun_out/s9_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s9_tests.log
tail -5 gpurun_out/s9_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/s9_bench_n2.json 2> gpurun_out/s9_bench_n2.err
echo "bench rc=$?"
tail -3 gpurun_out/s9_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s9_bench_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","n_gpus","ms_per_step","e2e","gpu_launches") if k in d})
print({k: v for k, v in d.items() if "gather" in k or "pipel" in k or "sync" in k})
PY
