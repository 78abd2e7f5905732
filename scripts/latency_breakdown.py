"""Where a single default-lattice replan's time goes (host perf_counter around the pieces of generate_range_polynomials)."""
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
def timeit(fn, n=400):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return np.median(ts) * 1e6
one = nat.STAGE_K1 | nat.STAGE_K2 | nat.STAGE_K5 | nat.STAGE_GATHER | nat.STAGE_ONE_LAUNCH
sep = one & ~nat.STAGE_ONE_LAUNCH
print("replanner total                   %.1f us" % timeit(lambda: pl.replanner(7.0)))
print("run_graph (one launch)            %.1f us" % timeit(lambda: eng.run_graph(P, None, one, 0, nat.SEL_CTOT)))
print("run_graph (separate launches)     %.1f us" % timeit(lambda: (eng.run_graph(P, None, sep, 0, nat.SEL_CTOT))))
print("eng.run, no graph (one launch)    %.1f us" % timeit(lambda: eng.run(P, None, one, 0, nat.SEL_CTOT)))
print("  upload only (async)             %.1f us" % timeit(lambda: eng.upload(record_event=False)))
torch.cuda.synchronize()
print("  launch only (async, one launch) %.1f us" % timeit(lambda: eng.launch(P, None, one, 0, nat.SEL_CTOT)))
torch.cuda.synchronize()
print("  download+sync (idle stream)     %.1f us" % timeit(lambda: eng.download(sync=True)))
print("  _frenet_from_winner             %.1f us" % timeit(lambda: pl._frenet_from_winner("argmin_ctot")))
print("  optimal_at_time x2              %.1f us" % timeit(lambda: (pl.optimal_at_time(7.0, pl.opt_path_tot, "d"), pl.optimal_at_time(7.0, pl.opt_path_tot, "s"))))
key = [k for k in eng._graphs if k[2] == one][0]
g = eng._graphs[key]
st = torch.cuda.current_stream()
def replay_only():
    g.replay()
t_replay = timeit(replay_only); torch.cuda.synchronize()
def replay_sync():
    g.replay(); st.synchronize()
print("  graph replay call alone         %.1f us" % t_replay)
print("  graph replay + sync             %.1f us" % timeit(replay_sync))
# device time of the graph: events around the replay
ts = []
for _ in range(100):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
print("  graph device time (events)      %.1f us" % np.median(ts))
# host staging of generate_range_polynomials without the launch
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(2000):
    pl.replanner(7.0)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
