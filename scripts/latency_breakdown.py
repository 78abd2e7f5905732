"""Where a single replan's ~70 us go (host perf_counter around the pieces of generate_range_polynomials)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import otg_b200  # noqa
from otg_b200 import _native as nat
from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
pl = TrajectoryPlannerV1(TrajectoryPlannerParams())
pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(0.0, 0.0, 0.0))
for i in range(50):
    pl.replanner(0.2 + 0.1 * i)
eng = pl._engine
P = pl._params
def timeit(fn, n=300):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return np.median(ts) * 1e6
stages = nat.STAGE_K1 | nat.STAGE_K2 | nat.STAGE_K5 | nat.STAGE_GATHER
print("replanner total        %.1f us" % timeit(lambda: pl.replanner(7.0)))
print("eng.run (h2d+4 kernels+d2h+sync) %.1f us" % timeit(lambda: eng.run(P, None, stages, 0, nat.SEL_CTOT)))
print("  upload only (async)  %.1f us" % timeit(lambda: eng.upload()))
torch.cuda.synchronize()
print("  launch only (async)  %.1f us" % timeit(lambda: eng.launch(P, None, stages, 0, nat.SEL_CTOT)))
torch.cuda.synchronize()
print("  download+sync        %.1f us" % timeit(lambda: eng.download(sync=True)))
print("  _frenet_from_winner  %.1f us" % timeit(lambda: pl._frenet_from_winner("argmin_ctot")))
from otg_b200.engine import make_params
from otg_b200 import lattice as lat
print("  make_params          %.1f us" % timeit(lambda: make_params(pl, pl._rule, 0.1, False)))
print("  optimal_at_time x2   %.1f us" % timeit(lambda: (pl.optimal_at_time(7.0, pl.opt_path_tot, "d"), pl.optimal_at_time(7.0, pl.opt_path_tot, "s"))))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    eng.run(P, None, stages, 0, nat.SEL_CTOT)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    eng.upload(); eng.launch(P, None, stages, 0, nat.SEL_CTOT); eng.download(sync=False)
def replay():
    g.replay(); torch.cuda.current_stream().synchronize()
print("graph replay + sync    %.1f us" % timeit(replay))
