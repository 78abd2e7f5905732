#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "one_launch" > gpurun_out/s37_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s37_tests.log
tail -15 gpurun_out/s37_tests.log
timeout 900 python - <<'PY' 2>&1 | tail -8
import sys, json
sys.path.insert(0, ".")
import torch, bench, numpy as np
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
wl = bench.workload(0, 512)
import otg_b200
from otg_b200 import workloads as w
from otg_b200.batch import BatchPlanner
from otg_b200.trajectory import SplineTrajectory2D
bp = BatchPlanner(bench.dense_params(), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), 512, bench.N_OBS, "v1", dev)
bp.stage_inputs(**wl); bp.engine.upload()
for _ in range(3): bp.launch_resident()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): bp.launch_resident()
b.record(); torch.cuda.synchronize()
print("three launches: %.3f ms per step" % (a.elapsed_time(b) / 20))
bp.engine.download(sync=True)
rec = np.array(bp.engine.result, copy=True)
del bp; torch.cuda.empty_cache()
print("config 5 one launch:", json.dumps(bench.one_launch_mode(wl, 20, dev, rec)))
print("default lattice batch:", json.dumps(bench.default_lattice_batch(dev)))
PY
