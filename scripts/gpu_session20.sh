#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PKG=optimal-trajectory-generation-in-the-duckietown-environment_b200
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s20_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s20_tests.log
tail -5 gpurun_out/s20_tests.log
OTG_B200_LIB=$PWD/$PKG/ab/libfrenet_phase.so timeout 300 python scripts/phase_times.py 2>&1 | grep -v Warn > gpurun_out/s20_phase_times.txt
cat gpurun_out/s20_phase_times.txt
timeout 300 python scripts/latency_modes.py > gpurun_out/s20_latency_modes.txt 2>&1
tail -3 gpurun_out/s20_latency_modes.txt
