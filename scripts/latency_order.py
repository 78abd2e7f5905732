"""Does the order of the latency legs matter?  default / no-speculation alternating, config 1 only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import otg_b200  # noqa
from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams


def config1(n_calls=300, **kw):
    pl = TrajectoryPlannerV1(TrajectoryPlannerParams(), **kw)
    pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(0.0, 0.0, 0.0))
    ts = []
    for i in range(n_calls + 20):
        t = time.perf_counter()
        pl.replanner(0.1 + 0.1 * (i + 1))
        ts.append(time.perf_counter() - t)
    ts = np.sort(np.array(ts[20:])) * 1e6
    return round(float(ts[len(ts) // 2]), 1), round(float(ts[len(ts) // 10]), 1)


for rep in range(3):
    print("default", config1(), "no_speculation", config1(_speculate=False), "default again", config1(), "inline", config1(_inline=True))
