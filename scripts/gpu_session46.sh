#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s46_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s46_tests.log
tail -8 gpurun_out/s46_tests.log
