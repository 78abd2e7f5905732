#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== full suite" > gpurun_out/s6_tests.log
timeout 1500 python -m pytest tests -m gpu -x -q >> gpurun_out/s6_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s6_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
echo "bench rc=$?" >> gpurun_out/s6_tests.log
timeout 900 python scripts/parity_sweep.py --field spread --out gpurun_out/r2_parity_sweep_spread > gpurun_out/s6_sweep_spread.log 2>&1
echo "sweep spread rc=$?" >> gpurun_out/s6_tests.log
timeout 900 python scripts/parity_sweep.py --field literal --out gpurun_out/r2_parity_sweep_literal > gpurun_out/s6_sweep_literal.log 2>&1
echo "sweep literal rc=$?" >> gpurun_out/s6_tests.log
tail -8 gpurun_out/s6_tests.log
tail -2 gpurun_out/s6_sweep_spread.log | cut -c1-600
tail -2 gpurun_out/s6_sweep_literal.log | cut -c1-600
