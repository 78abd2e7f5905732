"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel once, odd sizes.

    compute-sanitizer --tool memcheck python scripts/sanitize_case.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import otg_b200  # noqa: E402,F401
import _gpu  # noqa: E402
from otg_b200 import _native as nat  # noqa: E402
from otg_b200.batch import BatchPlanner  # noqa: E402
from otg_b200.planner import TrajectoryPlannerParams  # noqa: E402
from otg_b200.trajectory import CircleTrajectory2D, SplineTrajectory2D  # noqa: E402
from otg_b200 import workloads as w  # noqa: E402
from oracle import frenet_oracle as fo  # noqa: E402

rng = np.random.default_rng(3)
B = 5
p0 = np.stack([rng.uniform(-1, 1, B), rng.uniform(-.5, .5, B), rng.uniform(-.5, .5, B)], 1)
s0 = np.stack([rng.uniform(0, 25, B), rng.uniform(0, 4, B), rng.uniform(-.5, .5, B)], 1)
for variant in ("v1", "dt", "dt_obstacles", "optimal"):
    prm = fo.OracleParams.defaults(variant).with_(v_max=6.0, a_max=4.0, kappa_max=2.0)
    for tr in (CircleTrajectory2D(8, 5, 2), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY)):
        obs = rng.uniform(-5, 30, (B, 7, 2))
        for nf in (False, True):
            e = _gpu.run_batch(prm, p0, s0, 0.3, dd=0.2 if variant != "dt" else 0.1, trajectory=tr, obstacles=obs,
                               active=(rng.uniform(size=(B, 7)) < 0.8), use_mask=nat.MASK_KINEMATIC, curvature=nf, no_fuse=nf)
            assert (e.result["n_valid"] >= 0).all()
params = TrajectoryPlannerParams(d_road_width=0.5, dt=0.5, num_sample=2, delta_t=0.07)
bp = BatchPlanner(params, SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), B, 9, predict_steps=33)
r = bp.plan(p0, s0, np.full(B, 0.1), obs_s=s0[:, :1] + rng.uniform(2, 20, (B, 9)), obs_d=rng.uniform(-4, 4, (B, 9)),
            obs_speed=rng.uniform(0, 1, (B, 9)))
eng = bp.engine
x, y = eng.points_to_cartesian(bp.path, np.linspace(-1, 60, 77), np.linspace(-2, 2, 77))
free, first = eng.polyline_collision([x[:30], x[30:]], [y[:30], y[30:]], np.array([[x[3], y[3]], [100.0, 100.0]]), 0.49)
assert not free[0] and first[0] == 0
print("sanitize case ok", r.records["status"].tolist())
