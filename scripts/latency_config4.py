"""Where the ~210 us of a config-4 replan (dense lattice, one planner, 64 obstacles) go: host perf_counter around the three driver-level
calls, and CUDA-event times of the kernels behind them."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import otg_b200  # noqa: F401,E402
from otg_b200 import _native as nat, drivers, workloads as w  # noqa: E402
from otg_b200.drivers import moving_obstacle_planner as mv  # noqa: E402
from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams  # noqa: E402
from otg_b200.sensor import StaticObstacle  # noqa: E402
from otg_b200.trajectory import SplineTrajectory2D  # noqa: E402

sc = w.config4_scenario()
spl = SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY)
pl = TrajectoryPlannerV1(TrajectoryPlannerParams(**w.DENSE_PARAMS), _n_obs_cap=64)
pl.initialize(t0=sc["t0"], p0=sc["p0"], s0=sc["s0"], dd=sc["dd"])
eng = pl._engine
ox, oy = eng.points_to_cartesian(eng.path_handle(spl), sc["obs_s"], sc["obs_d"])


class AllSeeing:
    obstacles = [StaticObstacle(np.array([x, y])) for x, y in zip(ox, oy)]

    def sense(self, *a, **k):
        return self.obstacles


sensor = AllSeeing()
t_a, t_b, t_c = [], [], []
for i in range(120):
    pl.initialize(t0=sc["t0"], p0=sc["p0"], s0=sc["s0"], dd=sc["dd"]) if i == 0 else None
    t0 = time.perf_counter()
    pl.replanner(sc["t0"] + 0.02)
    t1 = time.perf_counter()
    paths = drivers.frenet_to_glob(pl, spl, pl.paths)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    pl.paths, _ = mv.check_paths(paths, sensor, None)
    pl.opt_path_tot = drivers.best_path(pl.paths)
    t3 = time.perf_counter()
    t_a.append(t1 - t0); t_b.append(t2 - t1); t_c.append(t3 - t2)
    pl.initialize(t0=sc["t0"], p0=sc["p0"], s0=sc["s0"], dd=sc["dd"])
med = lambda v: float(np.median(v[20:]) * 1e6)   # noqa: E731
print(f"replanner {med(t_a):.1f} us   frenet_to_glob + sync {med(t_b):.1f} us   check_paths + pick {med(t_c):.1f} us")
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L, P, Bf, path = C.byref(eng.L), C.byref(pl._params), C.byref(eng.B), C.byref(eng.path_handle(spl).desc)
calls = {"k1": lambda: nat.lib.frenet_k1_coefficients(L, P, Bf, stream), "k2": lambda: nat.lib.frenet_k2_evaluate(L, P, Bf, stream),
         "k3": lambda: nat.lib.frenet_k3_cartesian(L, P, path, Bf, stream), "k4": lambda: nat.lib.frenet_k4_collision(L, P, Bf, stream),
         "k5": lambda: nat.lib.frenet_k5_argmin(L, Bf, 4, stream), "gather": lambda: nat.lib.frenet_gather_winner(L, P, Bf, 4, 1, stream),
         "fused k2k3k4": lambda: nat.lib.frenet_k2_fused(L, P, path, Bf, 1, stream)}
for name, fn in calls.items():
    nat.check(fn())
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nat.check(fn()); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f"  {name:14s} {np.median(ts):7.1f} us")
