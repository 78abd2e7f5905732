#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "c_host" > gpurun_out/s47_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s47_tests.log
tail -25 gpurun_out/s47_tests.log
