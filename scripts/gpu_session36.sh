#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python scripts/parity_sweep.py --field spread --out gpurun_out/r2_parity_sweep_spread > gpurun_out/s36_sweep_spread.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/s36_sweep_spread.log
