#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/latency_breakdown.py > gpurun_out/s16_latency_breakdown.txt 2>&1
grep -v "^$" gpurun_out/s16_latency_breakdown.txt | head -45
