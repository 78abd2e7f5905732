"""A few default-lattice replans through the drop-in API (for ncu launch lists of the single-planner path)."""
import sys
sys.path.insert(0, ".")
import otg_b200  # noqa: F401,E402
from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams  # noqa: E402
kw = {} if len(sys.argv) < 3 else dict(_one_launch=False)
pl = TrajectoryPlannerV1(TrajectoryPlannerParams(), **kw)
pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(0.0, 0.0, 0.0))
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    pl.replanner(0.2 + 0.1 * i)
print(pl.opt_path_tot)
