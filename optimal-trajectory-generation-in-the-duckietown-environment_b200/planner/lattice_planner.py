"""The four planner classes of the reference, backed by the CUDA lattice engine.

Drop-in for `lib/planner` (`TrajectoryPlannerV1`, `TrajectoryPlannerV1DT`, `TrajectoryPlannerV1DTObstacles`,
`TrajectoryPlannerOptimal`): same constructors, methods, return values, attribute names and mutability.
What differs is where the work happens - `generate_range_polynomials` uploads the planner state and the
lattice axes, runs K1 (coefficients), K2 (sampling / costs) and K5 (argmins) on the GPU and brings back
one 64-byte record plus the winning trajectory.  The candidate list itself (`planner.paths`) stays in HBM
and is materialised into `Frenet` objects only when a caller looks at it.

There is no CPU implementation behind these classes: without a CUDA device they raise.
"""
from __future__ import annotations

import logging
from operator import attrgetter

import numpy as np

import struct

from .. import _native as nat
from .. import lattice as lat
from ..engine import FrenetEngine, LatticeGeometry, make_params
from .frenet import Frenet, WinnerFrenet
from .params import (TrajectoryPlannerParams, TrajectoryPlannerParamsDT, TrajectoryPlannerParamsDTObstacles,
                     TrajectoryPlannerParamsOptimal)
from .base import Planner

logger = logging.getLogger(__name__)

_WINNER_ROWS = ("t", "d", "dot_d", "ddot_d", "dddot_d", "s", "dot_s", "ddot_s", "dddot_s", "x", "y")
_SEL_OF_KEY = {"ctot": nat.SEL_CTOT, "cd": nat.SEL_CD, "cv": nat.SEL_CV, "ct": nat.SEL_CT}
_RESULT_OF_KEY = {"ctot": "argmin_ctot", "cd": "argmin_cd", "cv": "argmin_cv", "ct": "argmin_ct"}


class LatticePaths:
    """The candidate list of one replan, living in HBM.

    Behaves like the reference's `[Frenet, ...]`: `len`, indexing, iteration, `index`, `min(..., key=...)`.
    Elements are built on first access from one device-to-host copy of the scenario's arrays.  The view
    refers to the planner's device buffers, which the next replan overwrites: touching a view that has not
    been materialised after a newer replan raises RuntimeError (call `materialize()` first to keep it).
    """

    def __init__(self, planner, n_v: int, follow: bool, engine=None):
        self._planner = planner
        self._engine = engine if engine is not None else planner._engine
        self._generation = self._engine.generation
        self._n_v = n_v
        self._follow = follow
        self._fields = None
        self._valid_idx = None
        self._objs = {}
        self._t_initial = planner.t_initial
        self._n_all = n_v * self._engine.geom.n_t * self._engine.geom.n_d
        self._subset = None       # positions (into the valid list) kept by a filter, or None = all
        self._pending_mask = None  # callable -> bool mask over ALL candidates, resolved on first element access
        self._known_len = None    # length known from a K5 record without touching the candidate arrays
        self._best = {}           # cost key -> Frenet already gathered on the device (e.g. the masked winner)

    # -- device -> host ---------------------------------------------------------------------------
    def _check_live(self):
        if self._engine.generation != self._generation:
            raise RuntimeError("this candidate list belongs to an earlier replan whose device buffers were reused; "
                               "call .materialize() before replanning to keep it")

    def materialize(self):
        if self._fields is None:
            self._check_live()
            e = self._engine
            f = e.padded_fields(e.fetch_scenario(0))
            self._fields = f
            self._valid_idx = np.nonzero(f["valid"])[0]
        elif ("x" not in self._fields and self._engine.has_xy and self._engine.generation == self._generation):
            e = self._engine     # K3 ran after the first fetch: add x, y
            f = e.padded_fields(e.fetch_scenario(0))
            self._fields.update(x=f["x"], y=f["y"])
            for o in self._objs.values():
                n = len(o.t)
                o.x, o.y = f["x"][o.index, :n].tolist(), f["y"][o.index, :n].tolist()
        return self

    def _positions(self):
        self.materialize()
        if self._pending_mask is not None:
            keep = np.asarray(self._pending_mask(), dtype=bool)[self._valid_idx]
            base = np.arange(len(self._valid_idx)) if self._subset is None else self._subset
            self._subset = base[keep[base]]
            self._pending_mask = None
        return self._valid_idx if self._subset is None else self._valid_idx[self._subset]

    # -- sequence protocol -------------------------------------------------------------------------
    def __len__(self):
        if self._known_len is not None:
            return self._known_len
        if self._pending_mask is not None:
            return len(self._positions())
        if self._subset is not None:
            return len(self._subset)
        if self._fields is None and self._planner._rule != nat.LOWSPEED_OPTIMAL:
            return self._n_all     # no row can be dropped
        return len(self._positions())

    def _build(self, c: int) -> Frenet:
        o = self._objs.get(c)
        if o is None:
            f = self._fields
            n = int(f["nsteps"][c])
            o = Frenet()
            o.index = c
            period = self._engine.geom.period
            o.t = (np.arange(n) * period + self._t_initial).tolist()
            for name in _WINNER_ROWS[1:9]:
                setattr(o, name, f[name][c, :n].tolist())
            if "x" in f:
                o.x, o.y = f["x"][c, :n].tolist(), f["y"][c, :n].tolist()
            o.cd, o.cv, o.ct, o.ctot = (float(f[k][c]) for k in ("cd", "cv", "ct", "ctot"))
            self._objs[c] = o
        return o

    def __getitem__(self, i):
        pos = self._positions()
        if isinstance(i, slice):
            return [self._build(int(c)) for c in pos[i]]
        return self._build(int(pos[i]))

    def __iter__(self):
        for c in self._positions():
            yield self._build(int(c))

    def index(self, frenet) -> int:
        pos = self._positions()
        hit = np.nonzero(pos == frenet.index)[0]
        if hit.size == 0:
            raise ValueError("not in list")
        return int(hit[0])

    def filtered(self, keep_mask: np.ndarray) -> "LatticePaths":
        """A view of the elements of this view whose entry in `keep_mask` (length len(self)) is true; order preserved."""
        self.materialize()
        out = LatticePaths.__new__(LatticePaths)
        out.__dict__.update(self.__dict__)
        base = np.arange(len(self._valid_idx)) if self._subset is None else self._subset
        out._subset = base[np.asarray(keep_mask, dtype=bool)]
        out._known_len, out._best = None, {}
        return out

    def masked_view(self, mask_fn, known_len: int, best: dict) -> "LatticePaths":
        """A filtered view whose mask (over all candidates) is fetched from the device only if elements are
        looked at; its length and its cheapest element are already known from the K5 record."""
        out = LatticePaths.__new__(LatticePaths)
        out.__dict__.update(self.__dict__)
        out._pending_mask, out._known_len, out._best = mask_fn, int(known_len), dict(best)
        return out

    def argmin(self, key: str = "ctot"):
        """First minimal element by cost `key` (what `min(paths, key=attrgetter(key))` returns); ValueError if empty."""
        if key in self._best:
            if self._best[key] is None:
                raise ValueError("min() arg is an empty sequence")
            return self._best[key]
        pos = self._positions()
        if len(pos) == 0:
            raise ValueError("min() arg is an empty sequence")
        vals = self._fields[key][pos]
        return self._build(int(pos[int(np.argmin(vals))]))


_RESULT_I32 = {name: i for i, (name, *_rest) in enumerate(nat.RESULT_DTYPE[:8])}     # the int32 fields lead the 64-byte record
_pack_state = struct.Struct("<6d").pack_into
_pack_scal = struct.Struct("<3d").pack_into
_pack_int = struct.Struct("<i").pack_into


class _LatticePlanner(Planner):
    """Shared implementation; subclasses pick the variant rules (variant table)."""
    _variant = "v1"
    _params_cls = TrajectoryPlannerParams

    def __init__(self, *args, **kwargs):
        if args and isinstance(args[0], self._params_cls):
            self.__dict__.update(args[0].__dict__)
        if kwargs:
            self.__dict__.update(kwargs)

    # -- engine plumbing -----------------------------------------------------------------------------
    @property
    def _rule(self):
        return lat.VARIANTS[self._variant]["rule"]

    @property
    def _engine(self) -> FrenetEngine:
        e = self.__dict__.get("_engine_obj")
        if e is None:
            e = FrenetEngine(1, self.__dict__.get("device"), zero_copy=self.__dict__.get("_zero_copy", "out"), masked_winner=True)
            self.__dict__["_engine_obj"] = e
            self.__dict__["_geom_cache"] = {}
            self.__dict__["_n_v_cap"] = 0
            self.__dict__.setdefault("_n_obs_cap", 64)     # obstacle-table capacity of check_paths (constructor kwarg)
        return e

    def _sample_period(self):
        return lat.sample_period(self, self._variant)

    def step(self, t: float, dd: float = None, dsd: float = None, s_target=None):
        pass

    # -- reference API ---------------------------------------------------------------------------------
    def initialize(self, t0, p0, s0, dd: float = 0, dsd: float = 5, s_target=None):
        """lib/planner/trajectory_planner.py:86-106 (`dsd` is accepted and ignored, like the reference)."""
        assert abs(dd) <= self.max_road_width, 'desired offset must be into road limits'
        self.s0 = s0
        self.p0 = p0
        self.s_target = s_target
        self.dd = dd
        self.t_initial = t0
        self.d_target = 2.
        self.di_interval = lat.di_interval(self, self._variant)
        self.t_interval = lat.t_interval(self, self._variant)
        self.si_interval = lat.si_interval(self)
        self.dsi_interval = lat.dsi_interval(self, self.s0[1])
        self.paths = self.generate_range_polynomials()
        self.opt_path_d = self._winner("cd")
        self.opt_path_s = self._winner("cv")
        self.opt_path_tot = self._winner("ctot")
        if self.s_target is not None:
            self.opt_path_ct = self._winner("ct")
        self.opt_path_cd = self.opt_path_d
        self.opt_path_cv = self.opt_path_s

    def _target_triples(self, T: np.ndarray) -> np.ndarray:
        """(st, dst, ddst) per horizon - V1 family: trajectory_planner.py:132-136 (ObstacleTrajectory target)."""
        s0, ds0, dds0 = self.s0
        tg = self.s_target
        out = np.empty((len(T), 3))
        for j, tj in enumerate(T):
            time = self.t_initial + tj
            out[j, 0] = s0 + tg.compute_s(time) - self.d_target + self.delta_t * tg.compute_ds(time)
            out[j, 1] = ds0 + tg.compute_ds(time) - self.delta_t * tg.compute_dds(time)
            out[j, 2] = dds0 + tg.compute_dds(time)
        return out

    def generate_range_polynomials(self):
        """The lattice of one replan (trajectory_planner.py:108-181): K1 + K2 + K5 on the device.

        Returns the candidate list as a lazy `LatticePaths` view."""
        d = self.__dict__
        eng = d.get("_engine_obj") or self._engine
        ds0 = self.s0[1]
        self.si_interval = lat.si_interval(self)
        self.dsi_interval = lat.dsi_interval(self, ds0)
        follow = self.s_target is not None
        iv = self.si_interval if follow else self.dsi_interval
        V = np.arange(iv[0], iv[1], iv[2])
        period = self._sample_period()
        gkey = (tuple(self.di_interval), tuple(self.t_interval), period)
        geom = self._geom_cache.get(gkey)
        if geom is None:
            geom = LatticeGeometry.from_intervals(self.di_interval, self.t_interval, period)
            self._geom_cache[gkey] = geom
        if len(V) > self._n_v_cap or self._n_v_cap == 0:      # an empty axis (no candidates at all) still needs a valid layout
            d["_n_v_cap"] = max(1, len(V), lat.max_long_axis_len(self))
        sig = (geom, self._n_v_cap, self._n_obs_cap)
        if d.get("_conf_sig") != sig:            # (configure compares the same things, a few microseconds slower)
            eng.configure(geom, self._n_v_cap, self._n_obs_cap)
            d["_conf_sig"] = sig
        eng.host_block_free()
        # state (p0, s0), scalars and the axis length straight into the pinned staging block
        hv, offs = eng._in_host_bytes, eng._in_offs
        _pack_state(hv, offs["state"], *self.p0, *self.s0)
        _pack_scal(hv, offs["scal"], self.dd, self.desired_speed, self.t_initial)
        _pack_int(hv, offs["n_v"], len(V))
        eng.axis_v[0, :len(V)] = V
        if follow:
            eng.target[0] = self._target_triples(geom.T)
        # the parameter block is rebuilt only when one of the attributes it is made from changed (callers may mutate them)
        pkey = (self.kj, self.kt, self.kd, self.ks, self.kdots, self.klat, self.klong, self.low_speed_threshold, period, follow,
                d.get("S_thresh"), d.get("v_max"), d.get("a_max"), d.get("kappa_max"))
        if d.get("_params_key") != pkey:
            d["_params"] = make_params(self, self._rule, period, follow)
            d["_params_key"] = pkey
        P = self._params
        stages = nat.STAGE_K1 | nat.STAGE_K2 | nat.STAGE_K5 | nat.STAGE_GATHER
        if d.get("_one_launch", True):
            stages |= nat.STAGE_ONE_LAUNCH     # K1, the evaluate kernel and K5 + gather as ONE kernel (a replan is launch-bound)
        # the driver's next call is frenet_to_glob on the path it used last time (drivers.frenet_to_glob remembers it here): the
        # transform is fused into this launch and that call finds x, y already there
        spec = d.get("_spec_path")
        handle = None
        if spec is not None and not eng.relative_s:
            stages |= nat.STAGE_K3
            handle = eng.path_handle(spec)
        # ... and if the driver presented the SAME obstacles to check_paths twice in a row (a static scene), it will most likely do
        # so again: the collision stage (nearly free inside the fused kernel) and the masked pick ride along, and check_paths only
        # has to compare its obstacle list with the one that was tested (drivers.collision_filter)
        spec_obs = d.get("_spec_obs")
        use_mask, n_obs, eng.spec_pos = 0, None, None
        if (handle is not None and spec_obs is not None and (stages & nat.STAGE_ONE_LAUNCH) and eng.masked_winner
                and eng.partials is not None and 0 < len(spec_obs) <= eng.n_obs):
            n_obs = len(spec_obs)
            eng.obs_xy[0, :n_obs] = spec_obs
            eng.obs_active[0, :n_obs] = 1
            stages |= nat.STAGE_K4
            use_mask = nat.MASK_COLLISION
            eng.spec_pos = spec_obs
        if (stages & nat.STAGE_ONE_LAUNCH) and d.get("_inline", False):
            # one launch with its inputs as the kernel argument (frenet_replan_inline): no graph, no copy node.  Measured 1-1.5 us
            # SLOWER per replan than replaying the two-node graph below (the 14-argument foreign call and a plain launch cost more
            # than cudaGraphLaunch), so it is an option, not the default; it is the form a host without graphs would use.
            eng.run_inline(P, handle, stages, use_mask, nat.SEL_CTOT, n_obs=n_obs, follow=follow)
        else:
            eng.run_graph(P, handle, stages, use_mask, nat.SEL_CTOT, n_obs=n_obs)
        eng.xy_path = spec if handle is not None else None
        best = self._frenet_from_winner("argmin_ctot")
        d["_winner_cache"] = {"ctot": best}
        paths = LatticePaths(self, len(V), follow, eng)
        paths._best["ctot"] = best     # None when the lattice is empty
        return paths

    def _frenet_from_winner(self, result_field: str, masked_block: bool = False):
        """Frenet object from the engine's winner block (None if the selection is empty); masked_block: from the second block,
        which a replan launched with obstacles fills with the masked pick."""
        eng = self.__dict__.get("_engine_obj") or self._engine
        block, steps = (eng.winner_masked, eng.winner_masked_steps) if masked_block else (eng.winner, eng.winner_steps)
        n = int(steps[0])
        if n == 0:
            return None
        # a private copy of the block (the pinned one is rewritten by the next launch); its rows become lists on first use
        rows = block[0, :11, :n].copy()
        return WinnerFrenet(rows, 11 if eng.has_xy else 9, block[0, 11, :4].tolist(), eng.result_i32.item(_RESULT_I32[result_field]))

    def _winner(self, key: str):
        """`min(self.paths, key=attrgetter(key))` of the CURRENT lattice (first minimal element wins)."""
        cache = self.__dict__.setdefault("_winner_cache", {})
        if key in cache:
            if cache[key] is None:      # empty lattice (every row dropped / empty axis): the reference's min([]) raises
                raise ValueError("min() arg is an empty sequence")
            return cache[key]
        eng = self._engine
        eng.launch(self._params, None, nat.STAGE_GATHER, 0, _SEL_OF_KEY[key])
        eng.download(sync=True)
        f = self._frenet_from_winner(_RESULT_OF_KEY[key])
        if f is None:
            raise ValueError("min() arg is an empty sequence")
        cache[key] = f
        return f

    def optimal_at_time(self, time, opt_path, type_path):
        """trajectory_planner.py:183-195 (the reference prints a warning at both clamps; logged here)."""
        if isinstance(opt_path, WinnerFrenet):
            val = opt_path._value          # reads the winner block directly unless the field was turned into a list already
        else:
            def val(name, i):
                return getattr(opt_path, name)[i]
        t_first = val("t", 0)
        if time <= t_first:
            logger.debug('requested time %s is before the optimal path interval', time)
            index = 0
        elif time >= val("t", -1):
            logger.debug('requested time %s is after the optimal path interval', time)
            index = -1
        else:
            index = round((time - t_first) / self._sample_period_for_lookup())
        if type_path == "s":
            return (val("s", index), val("dot_s", index), val("ddot_s", index))
        return (val("d", index), val("dot_d", index), val("ddot_d", index))

    def _sample_period_for_lookup(self):
        return self.delta_t

    def replan_ct(self, time: float, s_target=None):
        self.p0 = self.optimal_at_time(time, self.opt_path_ct, "d")
        self.s0 = self.optimal_at_time(time, self.opt_path_ct, "s")
        self.t_initial = time
        if s_target is not None:
            self.s_target = s_target
        self.paths = self.generate_range_polynomials()
        self.opt_path_ct = self._winner("ct")

    def replan_ctot(self, time: float):
        path = self.opt_path_tot
        both = path.state_at(time, self._sample_period_for_lookup()) if type(path) is WinnerFrenet else None
        if both is not None:           # the pick of the last replan, untouched: both look-ups straight from its block
            self.p0, self.s0 = both
        else:
            self.p0 = self.optimal_at_time(time, path, "d")
            self.s0 = self.optimal_at_time(time, path, "s")
        self.t_initial = time
        self.paths = self.generate_range_polynomials()
        self.opt_path_tot = self._winner("ctot")

    def replan_cd_cv(self, time: float):
        self.p0 = self.optimal_at_time(time, self.opt_path_d, "d")
        self.s0 = self.optimal_at_time(time, self.opt_path_s, "s")
        self.t_initial = time
        self.paths = self.generate_range_polynomials()
        self.opt_path_d = self._winner("cd")
        self.opt_path_s = self._winner("cv")

    def _apply_replanner_args(self, dd, dsd, s_target):
        if s_target is not None:       # follow has priority (trajectory_planner.py:224-231)
            self.s_target = s_target
        else:
            if dd is not None:
                self.dd = dd
            if dsd is not None:
                self.desired_speed = dsd

    def replanner(self, time: float, dd: float = None, dsd: float = None, s_target=None):
        self._apply_replanner_args(dd, dsd, s_target)
        self.replan_ctot(time=time)
        return self.s0, self.p0

    def _hold(self, time):
        """No replan this step: follow the stored optimum (trajectory_planner_dt.py:232-233)."""
        self.p0 = self.optimal_at_time(time, self.opt_path_tot, "d")
        self.s0 = self.optimal_at_time(time, self.opt_path_tot, "s")


class TrajectoryPlannerV1(_LatticePlanner):
    """lib/planner/trajectory_planner.py:69-234."""
    _variant = "v1"
    _params_cls = TrajectoryPlannerParams


class TrajectoryPlannerV1DT(_LatticePlanner):
    """lib/planner/trajectory_planner_dt.py:66-234: horizons stepped by delta_t, no low-speed mode,
    replans only near the ends of the stored optimum."""
    _variant = "dt"
    _params_cls = TrajectoryPlannerParamsDT

    def replanner(self, time: float, dd: float = None, dsd: float = None, s_target=None):
        self._apply_replanner_args(dd, dsd, s_target)
        t = self.opt_path_tot.t
        if time <= t[0] or time >= t[-2]:
            self.replan_ctot(time=time)
        else:
            self._hold(time)
        return self.s0, self.p0


class TrajectoryPlannerV1DTObstacles(_LatticePlanner):
    """lib/planner/trajectory_planner_obstacles.py:67-243: samples every `sampling_t`, lateral axis includes +W,
    replans within the last 30 samples of the stored optimum or when forced."""
    _variant = "dt_obstacles"
    _params_cls = TrajectoryPlannerParamsDTObstacles

    def _sample_period_for_lookup(self):
        return self.sampling_t

    def replanner(self, time: float, dd: float = None, dsd: float = None, s_target=None, force=False):
        self._apply_replanner_args(dd, dsd, s_target)
        t = self.opt_path_tot.t
        if (time <= t[0] or time >= t[-30]) or force:
            self.replan_ctot(time=time)
        else:
            self._hold(time)
        return self.s0, self.p0


class TrajectoryPlannerOptimal(_LatticePlanner):
    """lib/planner/trajectory_planner_optimal.py:69-235: `s_target` is a sampled `Frenet`; at low speed a
    candidate is kept only if it travels more than `S_thresh`."""
    _variant = "optimal"
    _params_cls = TrajectoryPlannerParamsOptimal

    def _target_triples(self, T: np.ndarray) -> np.ndarray:
        tg = self.s_target     # trajectory_planner_optimal.py:130-133
        out = np.empty((len(T), 3))
        for j, tj in enumerate(T):
            k = round((self.t_initial + tj - tg.t[0]) / self.delta_t)
            out[j] = (tg.s[k], tg.dot_s[k], tg.ddot_s[k])
        return out

    def replanner(self, time: float, dd: float = None, dsd: float = None, s_target=None):
        self._apply_replanner_args(dd, dsd, s_target)
        self.replan_ctot(time=time)
        return self.p0[0], self.s0[0], self.opt_path_tot
