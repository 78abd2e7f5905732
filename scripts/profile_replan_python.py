"""cProfile of the Python side of a config-1 replan (where the ~14 us outside the graph replay go)."""
import cProfile
import pstats
import sys

sys.path.insert(0, ".")
import otg_b200  # noqa: F401,E402
from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams  # noqa: E402

pl = TrajectoryPlannerV1(TrajectoryPlannerParams())
pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(0.0, 0.0, 0.0))
for i in range(200):
    pl.replanner(0.2 + 0.1 * i)
pr = cProfile.Profile()
pr.enable()
for i in range(3000):
    pl.replanner(20.3 + 0.1 * i)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
