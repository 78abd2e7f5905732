#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PKG=optimal-trajectory-generation-in-the-duckietown-environment_b200
LIBS="${1:-default div64}"
timeout 900 python -m pytest tests -m gpu -q -x -k "row_collision or fused or full_size or one_launch or kat" > gpurun_out/s34_tests.log 2>&1
tail -3 gpurun_out/s34_tests.log
for rep in 1 2; do
for lib in $LIBS; do
  if [ $lib = default ]; then unset OTG_B200_LIB; else export OTG_B200_LIB=$PWD/$PKG/ab/libfrenet_$lib.so; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --kernels-only > gpurun_out/s34_ab_${lib}_$rep.json 2> gpurun_out/s34_ab_${lib}_$rep.err
done
done
unset OTG_B200_LIB
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s34_ab_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", open(f.replace(".json",".err")).read()[-600:]); continue
    print(f, "step %.3f" % d["ms_per_step"], {n: round(v["ms"], 3) for n, v in d["kernels"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
