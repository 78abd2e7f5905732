"""Runs the UNMODIFIED reference (oracle/_ref, see build_ref.py) on the hot path - TEST INFRASTRUCTURE.

Only bench.py's CPU legs and scripts under scripts/ use this module; the product never does.  It imports the reference's own
planner class and driver functions and calls them exactly as `lib/tests/test_moving_obstacle_planner.py:86-98` does for one
replan of one planner: `TrajectoryPlannerV1.initialize` (-> generate_range_polynomials + min), `frenet_to_glob`,
`check_paths` (-> check_collisions per candidate) and `min(paths, key=ctot)` over the survivors.

The reference's driver modules import matplotlib / jsonpickle / gym / pyglet at module level for plotting and the Duckietown
simulator; none of that is on the planner path, so permissive stub modules are injected before the import (as in
tests/golden/_refimport.py, which does the same for fixture generation in the build container).
"""
from __future__ import annotations

import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        mod = _Stub(self.__name__ + "." + name)
        setattr(self, name, mod)
        return mod

    def __call__(self, *args, **kwargs):
        return _Stub("stub_call")


_STUBBED = ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches", "matplotlib.transforms",
            "matplotlib.lines", "matplotlib.gridspec", "matplotlib.collections", "jsonpickle", "gym", "gym.spaces",
            "gym_duckietown", "gym_duckietown.envs", "gym_duckietown.simulator", "pyglet", "pyglet.window"]

_mods = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "lib", "planner"))


def load():
    """Import the reference package from oracle/_ref (once per process)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference is mounted")
    sys.dont_write_bytecode = True
    for name in _STUBBED:
        sys.modules.setdefault(name, _Stub(name))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import logging
    import warnings
    logging.disable(logging.CRITICAL)      # the reference logs every planner step at INFO
    warnings.filterwarnings("ignore", category=SyntaxWarning)   # its plotting modules carry '\o' / '\d' in ordinary string literals
    from lib.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
    from lib.trajectory import CircleTrajectory2D, SplineTrajectory2D
    from lib.sensor import StaticObstacle
    import lib.tests.test_moving_obstacle_planner as moving
    _mods = types.SimpleNamespace(TrajectoryPlannerV1=TrajectoryPlannerV1, TrajectoryPlannerParams=TrajectoryPlannerParams,
                                  CircleTrajectory2D=CircleTrajectory2D, SplineTrajectory2D=SplineTrajectory2D,
                                  StaticObstacle=StaticObstacle, moving=moving)
    return _mods


class _AllSeeing:
    """Stands in for ProximitySensor where the workload defines every obstacle as sensed (configs 4 and 5)."""

    def __init__(self, obstacles):
        self.obstacles = obstacles

    def sense(self, rpose=None):
        return self.obstacles


def replan_pipeline(params: dict, spline_wxwy, p0, s0, t0, dd, obs_sd, num_sample=None):
    """One replan of one planner through the reference's own code.  obs_sd: (s [O], d [O]) Frenet positions of the obstacles.
    num_sample: override of the planner's `num_sample` (0 = the centre target speed only: a bounded sample of the same lattice).
    Returns a dict with the work done (candidate-timesteps), the picks and the wall time."""
    m = load()
    import contextlib
    import io
    from operator import attrgetter
    t_start = time.perf_counter()
    kw = dict(params)
    if num_sample is not None:
        kw["num_sample"] = num_sample
    spl = m.SplineTrajectory2D(*spline_wxwy)
    planner = m.TrajectoryPlannerV1(m.TrajectoryPlannerParams(**kw))
    with contextlib.redirect_stdout(io.StringIO()):      # optimal_at_time prints a warning at both clamps
        planner.initialize(t0=float(t0), p0=tuple(float(v) for v in p0), s0=tuple(float(v) for v in s0), dd=float(dd))
    paths = planner.paths
    n_all = len(paths)
    cand_ts = sum(len(f.t) for f in paths)
    argmin = paths.index(planner.opt_path_tot)
    m.moving.frenet_to_glob(planner, spl, paths)
    obstacles = [m.StaticObstacle(spl.compute_pt(s) + m.moving.compute_ortogonal_vect(spl, s) * d) for s, d in zip(*obs_sd)]
    ident = {id(f): i for i, f in enumerate(paths)}
    survivors, leading = m.moving.check_paths(paths, _AllSeeing(obstacles), None)
    if survivors:
        best = min(survivors, key=attrgetter("ctot"))
        argmin_masked, min_masked = ident[id(best)], float(best.ctot)
    else:
        argmin_masked, min_masked = -1, float("inf")
    return dict(cand_ts=int(cand_ts), n_candidates=n_all, argmin=int(argmin), argmin_masked=int(argmin_masked),
                n_survivors=len(survivors), min_masked_ctot=min_masked, seconds=time.perf_counter() - t_start)


def default_lattice_replan(n_calls=5):
    """Metric 2 with the reference itself: one replan of the default lattice (320 candidates) on the circle path,
    (a) `replanner` only, (b) + frenet_to_glob + check_paths (3 obstacles) + masked pick.  Median seconds per call."""
    m = load()
    import contextlib
    import io
    from operator import attrgetter
    tr = m.CircleTrajectory2D(8, 5, 2)
    obstacles = [m.StaticObstacle(np.array(p)) for p in ([5.0, 0.3], [6.0, -1.0], [4.5, 2.5])]
    t1, t3 = [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(n_calls + 1):
            pl = m.TrajectoryPlannerV1(m.TrajectoryPlannerParams())
            pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(3.0 + 0.01 * i, 2.0, 0.0))
            t = time.perf_counter()
            pl.replanner(0.2)
            ta = time.perf_counter()
            paths = m.moving.frenet_to_glob(pl, tr, pl.paths)
            pl.paths, _ = m.moving.check_paths(paths, _AllSeeing(obstacles), None)
            pl.opt_path_tot = min(pl.paths, key=attrgetter("ctot"))
            tb = time.perf_counter()
            t1.append(ta - t)
            t3.append(tb - t)
    return float(np.median(t1[1:])), float(np.median(t3[1:]))
