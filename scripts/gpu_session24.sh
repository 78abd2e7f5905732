#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s24_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s24_tests.log
tail -4 gpurun_out/s24_tests.log
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu --kernels-only > gpurun_out/s24_bench_n1.json 2> gpurun_out/s24_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu > gpurun_out/s24_bench_n2.json 2> gpurun_out/s24_bench_n2.err
echo "n2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_plan_sharded.py > gpurun_out/s24_sharded.log 2>&1; tail -2 gpurun_out/s24_sharded.log
python - <<'PY'
import json
for n in (1, 2):
    d=json.loads(open(f"gpurun_out/s24_bench_n{n}.json").read().strip().splitlines()[-1])
    print(n, "value %.4g ms %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"])
    if d.get("per_rank"):
        print("   with   ", [round(x, 3) for x in d["per_rank"]["ms_per_step_with_gather"]])
        print("   without", [round(x, 3) for x in d["per_rank"]["ms_per_step_without_gather"]])
PY
