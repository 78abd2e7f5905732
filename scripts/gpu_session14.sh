#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/config4_once.py 6 > gpurun_out/s14_once.log 2>&1
echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s14_config4_launches.csv python scripts/config4_once.py 6 > gpurun_out/s14_ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_evaluate --launch-skip 6 -c 2 -o gpurun_out/s14_config4_k2 python scripts/config4_once.py 8 > gpurun_out/s14_ncu.log 2>&1
tail -3 gpurun_out/s14_ncu.log
grep -v "^==" gpurun_out/s14_config4_launches.csv | awk -F'","' '{print $5, $(NF-6), $(NF-5), $NF}' | cut -c1-150 | tail -24
