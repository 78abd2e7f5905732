#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PKG=optimal-trajectory-generation-in-the-duckietown-environment_b200
: > gpurun_out/s26_fp32_ab.txt
for rep in 1 2; do
for lib in default c4 c5; do
  if [ $lib = default ]; then unset OTG_B200_LIB; else export OTG_B200_LIB=$PWD/$PKG/ab/libfrenet_$lib.so; fi
  timeout 300 python scripts/fp32_time.py 2>/dev/null | tail -1 >> gpurun_out/s26_fp32_ab.txt
done
done
cat gpurun_out/s26_fp32_ab.txt
