#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/config1_once.py 6 > gpurun_out/s17_once.log 2>&1
echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s17_config1_launches.csv python scripts/config1_once.py 6 > gpurun_out/s17_ncu_l.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s17_config1_launches_sep.csv python scripts/config1_once.py 6 sep > gpurun_out/s17_ncu_l2.log 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/s17_config1_launches.csv", "gpurun_out/s17_config1_launches_sep.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    h=rows[0]; ix={k:i for i,k in enumerate(h)}
    for r in rows[-6:]:
        print(r[ix['Kernel Name']][:70], r[ix['Grid Size']], r[ix['Metric Value']])
    print()
PY
timeout 300 python scripts/latency_breakdown.py > gpurun_out/s17_latency_breakdown.txt 2>&1
grep -v "^$" gpurun_out/s17_latency_breakdown.txt | head -14
