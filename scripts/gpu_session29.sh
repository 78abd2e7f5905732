#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python scripts/fused_vs_separate_sweep.py > gpurun_out/r2_fused_vs_separate_sweep.md 2> gpurun_out/s29_sweep.err
echo "rc=$?"; cat gpurun_out/r2_fused_vs_separate_sweep.md; tail -5 gpurun_out/s29_sweep.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
