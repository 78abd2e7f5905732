#!/bin/bash
# Round-2 GPU session 1: parity of the row-table kernel, A/B of the kernel builds, one ncu capture.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PKG=optimal-trajectory-generation-in-the-duckietown-environment_b200
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_gpu.txt 2>&1
echo "== new tests first" > gpurun_out/s1_tests.log
timeout 900 python -m pytest tests -m gpu -x -q -k "row_collision or short_horizon_kernel or fused_kernel or config5 or dense_config4" >> gpurun_out/s1_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s1_tests.log
for lib in default old bulk split1 cta4; do
  if [ $lib = default ]; then unset OTG_B200_LIB; else export OTG_B200_LIB=$PWD/$PKG/ab/libfrenet_$lib.so; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --kernels-only > gpurun_out/s1_ab_$lib.json 2> gpurun_out/s1_ab_$lib.err
done
unset OTG_B200_LIB
echo "== full suite" >> gpurun_out/s1_tests.log
timeout 1200 python -m pytest tests -m gpu -x -q >> gpurun_out/s1_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s1_tests.log
for seed in 2026 7 99; do
  echo "== soak fuzz seed $seed max_b 90" >> gpurun_out/s1_tests.log
  OTG_FUZZ_SEED=$seed OTG_FUZZ_MAX_B=90 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k randomised >> gpurun_out/s1_tests.log 2>&1
  echo "rc=$?" >> gpurun_out/s1_tests.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_evaluate --launch-skip 3 -c 1 -o gpurun_out/s1_fused_table \
  python bench.py --steps 2 --warmup 3 --no-cpu --kernels-only > gpurun_out/s1_ncu.log 2>&1
tail -3 gpurun_out/s1_tests.log
for lib in default old bulk split1 cta4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s1_ab_$lib.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print("$lib", "step %.3f ms"%d["ms_per_step"], {n:round(v["ms"],3) for n,v in k.items()}, d["clocks"])
except Exception as e:
    print("$lib", "failed", e)
PY
done
