#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "one_launch" > gpurun_out/s12_tests.log 2>&1
echo "rc=$?" >> gpurun_out/s12_tests.log
tail -3 gpurun_out/s12_tests.log
timeout 300 python scripts/latency_config4.py > gpurun_out/s12_latency_config4.txt 2>&1
cat gpurun_out/s12_latency_config4.txt
timeout 300 python scripts/latency_modes.py > gpurun_out/s12_latency_modes.txt 2>&1
tail -3 gpurun_out/s12_latency_modes.txt
python - <<'PY' > gpurun_out/s12_profile_config4.txt 2>&1
import cProfile, pstats, sys, numpy as np
sys.path.insert(0, ".")
import bench, otg_b200
from otg_b200 import drivers, workloads as w
from otg_b200.drivers import moving_obstacle_planner as mv
from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
from otg_b200.sensor import StaticObstacle
from otg_b200.trajectory import SplineTrajectory2D
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
s = AllSeeing()
def loop(n):
    for i in range(n):
        pl.replanner(sc["t0"])
        paths = drivers.frenet_to_glob(pl, spl, pl.paths)
        pl.paths, _ = mv.check_paths(paths, s, None)
        pl.opt_path_tot = drivers.best_path(pl.paths)
loop(50)
pr = cProfile.Profile(); pr.enable(); loop(1000); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
PY
head -45 gpurun_out/s12_profile_config4.txt
