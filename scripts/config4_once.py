"""A few config-4 replans through the drop-in API (for ncu captures of the single-planner kernels)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import otg_b200  # noqa: F401,E402
from otg_b200 import drivers, workloads as w  # noqa: E402
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


for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    pl.replanner(sc["t0"])
    paths = drivers.frenet_to_glob(pl, spl, pl.paths)
    pl.paths, _ = mv.check_paths(paths, AllSeeing(), None)
    pl.opt_path_tot = drivers.best_path(pl.paths)
print("survivors", len(pl.paths))
