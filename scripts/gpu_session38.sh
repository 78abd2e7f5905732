#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s38_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s38_tests.log
tail -6 gpurun_out/s38_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
