#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PKG=optimal-trajectory-generation-in-the-duckietown-environment_b200
OTG_B200_LIB=$PWD/$PKG/ab/libfrenet_phase.so timeout 300 python scripts/phase_times.py > gpurun_out/s19_phase_times.txt 2>&1
cat gpurun_out/s19_phase_times.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem --format=csv
