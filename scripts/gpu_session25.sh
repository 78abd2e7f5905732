#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/fp32_once.py > gpurun_out/s25_once.log 2>&1; tail -1 gpurun_out/s25_once.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_evaluate --launch-skip 2 -c 1 -o gpurun_out/s25_fp32_k2k3 python scripts/fp32_once.py > gpurun_out/s25_ncu.log 2>&1
tail -2 gpurun_out/s25_ncu.log
