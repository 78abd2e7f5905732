"""Fused pipeline kernel time against the number of obstacles per scenario (config-5 geometry, first k obstacles of each
scene): separates the fixed cost of the collision stage from the per-obstacle-group and per-survivor cost."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import otg_b200  # noqa: F401,E402
from otg_b200 import _native as nat, workloads as w  # noqa: E402
from otg_b200.batch import BatchPlanner  # noqa: E402
from otg_b200.planner import TrajectoryPlannerParams  # noqa: E402
from otg_b200.trajectory import SplineTrajectory2D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sc = w.config5_scenarios(B)
dev = torch.device("cuda", 0)
for k in (0, 1, 8, 16, 32, 64, 96, 128):
    bp = BatchPlanner(TrajectoryPlannerParams(**w.DENSE_PARAMS), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), B, max(k, 1), "v1", dev)
    kk = max(k, 1)
    obs_s, obs_d = sc["obs_s"][:, :kk].copy(), sc["obs_d"][:, :kk].copy()
    if k == 0:
        obs_d[:] = 1e6          # one obstacle, nowhere near: the fixed cost of the stage
    bp.stage_inputs(p0=sc["p0"], s0=sc["s0"], t0=sc["t0"], dd=sc["dd"], obs_s=obs_s, obs_d=obs_d)
    bp.engine.upload()
    bp.launch_resident()
    torch.cuda.synchronize()
    e = bp.engine
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L, P, Bf, path = C.byref(e.L), C.byref(bp.params), C.byref(e.B), C.byref(bp.path.desc)
    res = {}
    for coll in (1, 0):
        ts = []
        for _ in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nat.check(nat.lib.frenet_k2_fused(L, P, path, Bf, coll, stream))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[coll] = float(np.median(ts[2:]))
    e.download(sync=True)
    free = float(e.result["n_survivors"].sum()) / max(1.0, float(e.result["n_valid"].sum()))
    print(f"obstacles {k:4d}: fused+collision {res[1]:.3f} ms   fused {res[0]:.3f} ms   collision-free fraction {free:.3f}", flush=True)
    del bp
    torch.cuda.empty_cache()
