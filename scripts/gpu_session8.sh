#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s8_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s8_tests.log
tail -60 gpurun_out/s8_tests.log
