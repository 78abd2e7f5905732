#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python scripts/parity_sweep.py --field literal --out gpurun_out/r2_parity_sweep_literal > gpurun_out/s41_sweep_literal.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/s41_sweep_literal.log
