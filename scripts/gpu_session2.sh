#!/bin/bash
# Round-2 GPU session 2: interval-table kernel: parity, A/B of the kernel builds, full bench line, reference arm, ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PKG=optimal-trajectory-generation-in-the-duckietown-environment_b200
echo "== new tests first" > gpurun_out/s2_tests.log
timeout 900 python -m pytest tests -m gpu -x -q -k "row_collision or short_horizon_kernel or fused_kernel or config5 or dense_config4" >> gpurun_out/s2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s2_tests.log
for lib in default old bulk split1 split2 cta4; do
  if [ $lib = default ]; then unset OTG_B200_LIB; else export OTG_B200_LIB=$PWD/$PKG/ab/libfrenet_$lib.so; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --kernels-only > gpurun_out/s2_ab_$lib.json 2> gpurun_out/s2_ab_$lib.err
done
unset OTG_B200_LIB
echo "== full suite" >> gpurun_out/s2_tests.log
timeout 1200 python -m pytest tests -m gpu -x -q >> gpurun_out/s2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s2_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
echo "bench rc=$?" >> gpurun_out/s2_tests.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s2_bench_ref.json 2> gpurun_out/s2_bench_ref.err
echo "ref rc=$?" >> gpurun_out/s2_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_evaluate --launch-skip 3 -c 1 -o gpurun_out/s2_fused_table \
  python bench.py --steps 2 --warmup 3 --no-cpu --kernels-only > gpurun_out/s2_ncu.log 2>&1
tail -4 gpurun_out/s2_tests.log
for lib in default old bulk split1 split2 cta4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s2_ab_$lib.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print("$lib", "step %.3f ms"%d["ms_per_step"], {n:round(v["ms"],3) for n,v in k.items()}, d["clocks"])
except Exception as e:
    print("$lib", "failed", e)
PY
done
tail -c 600 gpurun_out/s2_bench.err
