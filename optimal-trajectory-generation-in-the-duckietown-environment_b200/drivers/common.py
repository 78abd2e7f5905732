"""Driver-level pieces of the hot path: Frenet -> Cartesian (K3), collision filter (K4), masked argmin (K5).

In the reference these live in the experiment modules, not in the planner:
`frenet_to_glob`, `check_paths`, `check_collisions`, `compute_ortogonal_vect` and the constant
`robot_radius` of lib/tests/test_obstacle_planner.py:16-56 and lib/tests/test_moving_obstacle_planner.py:15-57.
The functions here keep those names and signatures.  When they are handed the planner's own candidate list
(the `LatticePaths` view returned by `replanner`) they launch K3 / K4 on the lattice that is already in HBM;
handed any other list of Frenet-like objects they pack it and run the generic point / polyline kernels.
"""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from ..engine import FrenetEngine
from ..planner.lattice_planner import LatticePaths

robot_radius = 0.7

_generic_engine = None


def _engine_for_lists() -> FrenetEngine:
    global _generic_engine
    if _generic_engine is None:
        _generic_engine = FrenetEngine(1)
    return _generic_engine


def compute_ortogonal_vect(traj, s):
    """Unit normal of the reference path at arc length s (test_obstacle_planner.py:53-56): one point, host side."""
    g = traj.compute_first_derivative(s)
    th = np.arctan2(g[1], g[0])
    return np.array([-np.sin(th), np.cos(th)])


def _is_live(paths) -> bool:
    return isinstance(paths, LatticePaths) and paths._engine.generation == paths._generation


def _attach_xy_source(eng, frenet_paths, planner):
    """Picks that were gathered before K3 ran have no x, y; the reference's frenet_to_glob appends them to the same objects
    (`opt_path_tot` IS an element of `paths`).  They get a fetcher that reads the candidate's x / y block from the lattice the
    first time somebody looks (nothing on the replan path pays for it)."""
    from ..planner.frenet import WinnerFrenet
    gen = eng.generation
    n_v = int(eng.n_v[0])

    def source_for(w):
        def fetch(name):
            if eng.generation != gen or not eng.has_xy:
                return []
            lat_off, _, npad, nsteps = eng.cand_offsets(n_v)
            c = w.index
            j = 4 if name == "x" else 5
            at = int(lat_off[c]) + j * int(npad[c])
            return eng.lat[0, at:at + int(nsteps[c])].cpu().numpy().tolist()
        return fetch

    for cache in (frenet_paths._best, planner.__dict__.get("_winner_cache", {})):
        for w in cache.values():
            if isinstance(w, WinnerFrenet) and w._n_valid == 9 and "_xy_source" not in w.__dict__:
                w._xy_source = source_for(w)


def frenet_to_glob(planner, trajectory, frenet_paths):
    """Adds Cartesian x, y to every path (test_obstacle_planner.py:18-25): K3."""
    if _is_live(frenet_paths):
        eng = frenet_paths._engine
        eng.set_relative_s(False)
        if not eng.has_xy or getattr(eng, "xy_path", None) is not trajectory:
            eng.launch(planner._params, eng.path_handle(trajectory), nat.STAGE_K3)
            eng.xy_path = trajectory
            _attach_xy_source(eng, frenet_paths, planner)
        if hasattr(planner, "generate_range_polynomials") and planner.__dict__.get("_speculate", True):
            planner.__dict__["_spec_path"] = trajectory      # the next replan transforms against this path right away
        return frenet_paths
    eng = _engine_for_lists()
    paths = list(frenet_paths)
    if paths:
        s = np.concatenate([np.asarray(p.s, dtype=np.float64) for p in paths])
        d = np.concatenate([np.asarray(p.d, dtype=np.float64) for p in paths])
        x, y = eng.points_to_cartesian(eng.path_handle(trajectory), s, d)
        at = 0
        for p in paths:
            n = len(p.s)
            p.x.extend(x[at:at + n].tolist())     # the reference appends
            p.y.extend(y[at:at + n].tolist())
            at += n
    return frenet_paths


class LazyArray:
    """An integer array that is fetched from the device the first time somebody looks at it; its length is known up front."""

    def __init__(self, fetch, n):
        self._fetch, self._n, self._data = fetch, int(n), None

    def _get(self):
        if self._data is None:
            self._data = np.asarray(self._fetch())
        return self._data

    def __len__(self):
        return self._n

    def __iter__(self):
        return iter(self._get())

    def __getitem__(self, i):
        return self._get()[i]


class LeadingObstacles:
    """`leading_obstacles` of the moving-obstacle `check_paths`: one blocking obstacle per blocked candidate, resolved lazily."""

    def __init__(self, sensed, first):
        self._sensed, self._first = sensed, first

    def __len__(self):
        return len(self._first)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._sensed[int(j)] for j in np.asarray(self._first[:])[i]]
        return self._sensed[int(self._first[i])]

    def __iter__(self):
        return (self._sensed[int(j)] for j in self._first)


def collision_filter(frenet_paths, sensor, rpose, want_blockers: bool = False):
    """Shared body of both `check_paths` flavours.

    Returns (surviving paths, sensed obstacles, first-blocker index into the sensed list per BLOCKED path).
    `rpose` is passed to `sensor.sense` positionally like the reference does, which means the sensor uses its
    attached robot's pose (proximity_sensor.py:53-57)."""
    sensed = sensor.sense(rpose)
    if not sensed:
        if _is_live(frenet_paths):
            frenet_paths._planner.__dict__["_spec_obs"] = frenet_paths._planner.__dict__["_last_obs"] = None
        return frenet_paths, sensed, np.zeros(0, dtype=np.int64)
    if _is_live(frenet_paths):
        eng = frenet_paths._engine
        planner = frenet_paths._planner
        n = len(sensed)
        if n > eng.n_obs:
            # more obstacles than the planner's staging capacity (constructor kwarg `_n_obs_cap`, default 64): test the
            # lattice's Cartesian polylines with the generic kernel instead of re-allocating the lattice buffers
            frenet_paths.materialize()
            pos = frenet_paths._positions()
            f = frenet_paths._fields
            nst = f["nsteps"]
            free, first = eng.polyline_collision([f["x"][c, :nst[c]] for c in pos], [f["y"][c, :nst[c]] for c in pos],
                                                 np.array([ob.position for ob in sensed], dtype=np.float64), robot_radius ** 2)
            return frenet_paths.filtered(free), sensed, first[~free]
        pos = np.array([ob.position for ob in sensed], dtype=np.float64)
        # The replan speculates on the obstacle set when it repeated (a static scene): if it already tested exactly these positions,
        # masks, first blockers, the record and the masked pick are on hand and nothing is launched.
        last = planner.__dict__.get("_last_obs")
        planner.__dict__["_spec_obs"] = pos if (last is not None and last.shape == pos.shape and np.array_equal(last, pos)
                                                and planner.__dict__.get("_speculate", True)) else None
        planner.__dict__["_last_obs"] = pos
        tested = getattr(eng, "spec_pos", None)
        hit = tested is not None and eng.has_collision and tested.shape == pos.shape and np.array_equal(tested, pos)
        # (zero-copy engines: a K3 launch of frenet_to_glob may still be running - it never reads the obstacle arrays)
        if not hit:
            eng.spec_pos = None
            eng.obs_xy[0, :n] = pos
            eng.obs_active[0, :n] = 1
        # K4, then K5 with the collision mask and a gather of the masked winner: one graph replay (H2D of the obstacle
        # table, three kernels, D2H of the record + winner)
        xy_path = getattr(eng, "xy_path", None)
        if hit:
            pass
        elif (eng.partials is not None and planner.__dict__.get("_one_launch", True) and eng.has_xy
              and (eng.geom.n_pad_max <= 32 or (xy_path is not None and not eng.relative_s))):
            # the collision stage with the masked argmin + gather in its tail - ONE kernel instead of K4, then K5.  Long horizons: from
            # the row-level table (nothing is read back from the lattice; needs the path the x, y were computed against); short ones:
            # the sweep over the stored x, y
            long_h = eng.geom.n_pad_max > 32
            eng.run_graph(planner._params, eng.path_handle(xy_path) if long_h else None,
                          nat.STAGE_K4 | nat.STAGE_K5 | nat.STAGE_GATHER | nat.STAGE_ONE_LAUNCH, nat.MASK_COLLISION, nat.SEL_MASKED, n_obs=n)
        else:
            eng.run_graph(planner._params, None, nat.STAGE_K4 | nat.STAGE_K5 | nat.STAGE_GATHER, nat.MASK_COLLISION,
                          nat.SEL_MASKED, n_obs=n)
        C_ = frenet_paths._n_all
        winner = planner._frenet_from_winner("argmin_masked", masked_block=hit)

        def fetch_mask():
            frenet_paths._check_live()
            return eng.coll_free[0, :C_].cpu().numpy() != 0

        view = frenet_paths.masked_view(fetch_mask, int(eng.result[0]["n_survivors"]), {"ctot": winner})
        first = np.zeros(0, dtype=np.int64)
        if want_blockers:
            gen = eng.generation

            def fetch_blockers():      # only the all-blocked fallback of the moving-obstacle driver ever looks at these
                if eng.generation != gen:
                    raise RuntimeError("the blocking obstacles of an earlier replan are no longer on the device")
                free = eng.coll_free[0, :C_].cpu().numpy() != 0
                valid = (eng.cand_flags[0, :C_].cpu().numpy() & nat.CAND_VALID) != 0
                return eng.first_blocker[0, :C_].cpu().numpy()[valid & ~free]

            first = LazyArray(fetch_blockers, int(eng.result[0]["n_valid"]) - int(eng.result[0]["n_survivors"]))
        return view, sensed, first
    paths = list(frenet_paths)
    eng = _engine_for_lists()
    free, first = eng.polyline_collision([p.x for p in paths], [p.y for p in paths],
                                         np.array([ob.position for ob in sensed], dtype=np.float64), robot_radius ** 2)
    return [p for p, ok in zip(paths, free) if ok], sensed, first[~free]


def single_path_collision(path, obstacles):
    """(collision_free, first blocking obstacle or None) for one path: one launch of the polyline kernel."""
    if not obstacles:
        return True, None
    eng = _engine_for_lists()
    free, first = eng.polyline_collision([path.x], [path.y], np.array([ob.position for ob in obstacles], dtype=np.float64),
                                         robot_radius ** 2)
    return bool(free[0]), (None if free[0] else obstacles[int(first[0])])


def best_path(paths, key: str = "ctot"):
    """`min(paths, key=attrgetter(key))`: first minimal element; ValueError when `paths` is empty."""
    if isinstance(paths, LatticePaths):
        return paths.argmin(key)
    from operator import attrgetter
    return min(paths, key=attrgetter(key))
