"""Phase marks inside the one-launch replan kernel (debug build -DFRENET_PHASE_TIMING=1, selected with OTG_B200_LIB):
latest arrival of any CTA at each mark, relative to the first CTA's entry, median over replans.

    OTG_B200_LIB=$PWD/<pkg>/ab/libfrenet_phase.so python scripts/phase_times.py
"""
import ctypes as C
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import otg_b200  # noqa
from otg_b200 import _native as nat, drivers, workloads as w
from otg_b200.drivers import moving_obstacle_planner as mv
from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
from otg_b200.sensor import StaticObstacle
from otg_b200.trajectory import SplineTrajectory2D

raw = C.CDLL(nat.LIB_PATH)
names = ["entry (first CTA)", "inline inputs spilled", "K1 head done", "phase 0 staged", "phase A done", "phase B done", "results written",
         "partial reduced", "fence + ticket", "last CTA: partials combined", "record + winner written",
         "long kernel: warp 0's table share built", "long kernel: table resolved", "long kernel: exact re-tests done"]
order = [0, 1, 2, 3, 4, 11, 5, 12, 13, 6, 7, 8, 9, 10]


def marks():
    buf = (C.c_ulonglong * 16)()
    raw.frenet_debug_phase_times(buf, 1)
    return np.array(buf[:14], dtype=np.float64)


def report(title, fn, n=200):
    fn(); marks()
    rows, wall = [], []
    for _ in range(n):
        t = time.perf_counter(); fn(); wall.append(time.perf_counter() - t)
        m = marks()
        rows.append(np.where(m > 0, m - m[0], np.nan))
    med = np.nanmedian(np.array(rows), axis=0) / 1e3
    print(f"{title}: host wall {np.median(wall) * 1e6:.1f} us")
    for i in order:
        if not np.isnan(med[i]):
            print(f"   {i:2d} {names[i]:42s} {med[i]:7.2f} us")


pl = TrajectoryPlannerV1(TrajectoryPlannerParams())
pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(0.0, 0.0, 0.0))
report("config 1 replanner (k2_short, 24 CTAs)", lambda: pl.replanner(7.0))

sc = w.config4_scenario()
spl = SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY)
p4 = TrajectoryPlannerV1(TrajectoryPlannerParams(**w.DENSE_PARAMS), _n_obs_cap=64)
p4.initialize(t0=sc["t0"], p0=sc["p0"], s0=sc["s0"], dd=sc["dd"])
eng = p4._engine
ox, oy = eng.points_to_cartesian(eng.path_handle(spl), sc["obs_s"], sc["obs_d"])


class AllSeeing:
    obstacles = [StaticObstacle(np.array([x, y])) for x, y in zip(ox, oy)]

    def sense(self, *a, **k):
        return self.obstacles


def loop4():
    p4.replanner(sc["t0"])
    paths = drivers.frenet_to_glob(p4, spl, p4.paths)
    p4.paths, _ = mv.check_paths(paths, AllSeeing(), None)
    p4.opt_path_tot = drivers.best_path(p4.paths)


for _ in range(5):
    loop4()
report("config 4 loop (speculating: one fused kernel, K1 + K2 + K3 + K4 + K5)", loop4, 100)
