#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/s45_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s45_tests.log
tail -12 gpurun_out/s45_tests.log
timeout 300 python scripts/latency_modes.py > gpurun_out/s45_latency_modes.txt 2>&1
tail -3 gpurun_out/s45_latency_modes.txt
