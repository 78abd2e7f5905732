#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PKG=optimal-trajectory-generation-in-the-duckietown-environment_b200
echo "== table tests" > gpurun_out/s5_tests.log
timeout 900 python -m pytest tests -m gpu -x -q -k "row_collision or fused_kernel or config5 or dense_config4 or odd_sizes or error_codes or follow or closed_loop or replan_chain" >> gpurun_out/s5_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s5_tests.log
for lib in default old split1 split2 cta3 cta5; do
  if [ $lib = default ]; then unset OTG_B200_LIB; else export OTG_B200_LIB=$PWD/$PKG/ab/libfrenet_$lib.so; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --kernels-only > gpurun_out/s5_ab_$lib.json 2> gpurun_out/s5_ab_$lib.err
done
unset OTG_B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_evaluate --launch-skip 3 -c 1 -o gpurun_out/s5_fused_v2c4 \
  python bench.py --steps 2 --warmup 3 --no-cpu --kernels-only > gpurun_out/s5_ncu.log 2>&1
unset OTG_B200_LIB
tail -3 gpurun_out/s5_tests.log
for lib in default old split1 split2 cta3 cta5; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5_ab_$lib.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print("$lib", "step %.3f ms"%d["ms_per_step"], {n:round(v["ms"],3) for n,v in k.items() if n in ("k2k3k4_fused","k2k3_fused","k2_evaluate")}, d["clocks"])
except Exception as e:
    print("$lib", "failed", e)
PY
done
