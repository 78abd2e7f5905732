#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --kernels-only > gpurun_out/s30_bench_n1.json 2> gpurun_out/s30_bench_n1.err
for mode in equal balanced; do
  extra=""; [ $mode = balanced ] && extra="--balance-shards"
  SECONDS=0; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu $extra > gpurun_out/s30_bench_n${N}_$mode.json 2> gpurun_out/s30_bench_n${N}_$mode.err
  echo "$mode rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/s30_bench_n${N}_$mode.err
done
python - <<PY
import json
def show(f):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "failed", e); return
    print(f, "value %.4g ms %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["survivors"])
    if d.get("per_rank"):
        for k, v in d["per_rank"].items():
            if isinstance(v, list): print("   ", k, [round(x, 3) for x in v])
show("gpurun_out/s30_bench_n1.json")
show("gpurun_out/s30_bench_n${N}_balanced.json")
show("gpurun_out/s30_bench_n${N}_equal.json")
PY
