#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu --kernels-only > gpurun_out/s23_bench_n1.json 2> gpurun_out/s23_bench_n1.err
echo "n1 rc=$?"
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu > gpurun_out/s23_bench_n$n.json 2> gpurun_out/s23_bench_n$n.err
echo "n$n rc=$?"
done
python - <<'PY'
import json
for n in (1, 4, 8):
    try:
        d=json.loads(open(f"gpurun_out/s23_bench_n{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "failed", e); continue
    print(n, "value %.4g ms %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"])
    if d.get("per_rank"):
        print("   with   ", [round(x, 3) for x in d["per_rank"]["ms_per_step_with_gather"]])
        print("   without", [round(x, 3) for x in d["per_rank"]["ms_per_step_without_gather"]])
PY
