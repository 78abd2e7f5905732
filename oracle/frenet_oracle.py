"""CPU oracle for the Frenet lattice planner hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy fp64 restatement of the reference's algorithm; it is the checker that
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
run the CUDA path against.  Nothing in the product package imports it, and the product
path has no CPU fallback.

Parity pinning: the reference ships NO golden vectors or asserts for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the UNMODIFIED reference
run in the build container: `tests/golden/make_golden.py` imports `/root/reference` and
freezes its outputs into `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every
function below against those fixtures.  Where the reference calls into NumPy/LAPACK
(`np.linalg.solve`, `np.arange`, `**`) the oracle calls the same NumPy functions.  The
feasibility limits (v/a/kappa) and the time-indexed ("predicted") obstacle table have no
reference implementation at all: for those two extensions parity is UNPINNED and the
oracle is the definition.

Each function cites the reference file:line it follows (paths relative to /root/reference).
Conventions: rows r = (i_v, i_T) own one longitudinal polynomial, candidates
c = (i_v * n_T + i_T) * n_D + i_d own one lateral polynomial; N(T) = len(arange(0, T + dts, dts)).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace

import numpy as np

VARIANT_V1 = "v1"              # lib/planner/trajectory_planner.py
VARIANT_DT = "dt"              # lib/planner/trajectory_planner_dt.py
VARIANT_DTOBS = "dt_obstacles"  # lib/planner/trajectory_planner_obstacles.py
VARIANT_OPTIMAL = "optimal"    # lib/planner/trajectory_planner_optimal.py

ROBOT_RADIUS = 0.7             # lib/tests/test_obstacle_planner.py:16


@dataclass
class OracleParams:
    """Planner parameters (lib/planner/trajectory_planner.py:14-67 and the three variant files)."""
    variant: str = VARIANT_V1
    dt: float = 0.5
    kj: float = 0.001
    ks: float = 0.5
    kd: float = 0.5
    kt: float = 0.01
    kdots: float = 1
    klong: float = 1
    klat: float = 1
    delta_t: float = 0.1
    sampling_t: float = None      # DTObstacles only
    desired_speed: float = 2.5
    max_road_width: float = 4
    min_t: float = 1
    max_t: float = 3
    d_road_width: float = 0.5
    d_d_s: float = 1
    low_speed_threshold: float = 2
    num_sample: int = 2
    S_thresh: float = 2           # Optimal only
    # Extensions (no reference implementation; +inf == reference behaviour).
    v_max: float = math.inf
    a_max: float = math.inf
    kappa_max: float = math.inf

    @staticmethod
    def defaults(variant: str) -> "OracleParams":
        if variant == VARIANT_V1:
            return OracleParams()
        if variant == VARIANT_DT:  # trajectory_planner_dt.py:14-32
            return OracleParams(variant=VARIANT_DT, dt=None, ks=0.01, kd=1.0, desired_speed=0.5,
                                max_road_width=0.4, max_t=2, d_road_width=0.1, d_d_s=0.5,
                                low_speed_threshold=0.2)
        if variant == VARIANT_DTOBS:  # trajectory_planner_obstacles.py:14-33
            return OracleParams(variant=VARIANT_DTOBS, dt=None, ks=0.01, kd=2.0, klat=2, sampling_t=0.05,
                                delta_t=0.5, desired_speed=0.5, max_road_width=0.5, min_t=2, max_t=3,
                                d_road_width=0.1, d_d_s=0.2, low_speed_threshold=0.2)
        if variant == VARIANT_OPTIMAL:  # trajectory_planner_optimal.py:14-33
            return OracleParams(variant=VARIANT_OPTIMAL, dt=0.8, ks=0.1, desired_speed=2.0, max_road_width=2.45,
                                min_t=0.9, max_t=5.0, d_road_width=0.6, low_speed_threshold=2.0, num_sample=3)
        raise ValueError(variant)

    def with_(self, **kw) -> "OracleParams":
        return replace(self, **kw)

    # --- lattice intervals as `initialize` builds them -------------------------------------
    def di_interval(self):
        W, dd = self.max_road_width, self.d_road_width
        if self.variant == VARIANT_DTOBS:       # trajectory_planner_obstacles.py:92
            return (-W, W + dd, dd)
        return (-W, W, dd)                      # trajectory_planner.py:94

    def t_interval(self):
        if self.variant in (VARIANT_DT, VARIANT_DTOBS):   # trajectory_planner_dt.py:92
            return (self.min_t, self.max_t + self.delta_t, self.delta_t)
        return (self.min_t, self.max_t, self.dt)           # trajectory_planner.py:95

    def sample_period(self):
        return self.sampling_t if self.variant == VARIANT_DTOBS else self.delta_t


def long_axis(prm: OracleParams, ds0: float, follow: bool) -> np.ndarray:
    """Target-speed (velocity keeping) or target-offset (follow) axis.
    trajectory_planner.py:119-126 - Python banker's `round` on both ends, then np.arange."""
    n, q = prm.num_sample, prm.d_d_s
    if follow:
        iv = (round(-q * n), round(+q * n + q), q)
    else:
        iv = (round(ds0 - q * n), round(ds0 + q * n + q), q)
    return np.arange(iv[0], iv[1], iv[2])


def sample_times(T: float, period: float) -> np.ndarray:
    """trajectory_planner.py:139/148 - `np.arange(0, tj + delta_t, delta_t)`; may overshoot T."""
    return np.arange(0, T + period, period)


def quintic_coeffs(p0, dp0, ddp0, p1, dp1, ddp1, T):
    """lib/trajectory/trajectory_1d.py:15-40.  p1 may be an array (one system per entry)."""
    a0, a1, a2 = p0, dp0, ddp0 / 2
    A = np.array([[T ** 3, T ** 4, T ** 5],
                  [3 * T ** 2, 4 * T ** 3, 5 * T ** 4],
                  [6 * T, 12 * T ** 2, 20 * T ** 3]])
    p1 = np.atleast_1d(np.asarray(p1, dtype=np.float64))
    out = np.empty((p1.shape[0], 6))
    for i, p1_i in enumerate(p1):
        b = np.array([p1_i - a0 - a1 * T - a2 * T ** 2, dp1 - a1 - 2 * a2 * T, ddp1 - 2 * a2])
        out[i, 3:] = np.linalg.solve(A, b)
    out[:, 0], out[:, 1], out[:, 2] = a0, a1, a2
    return out


def quartic_coeffs(s0, ds0, dds0, ds1, dds1, T):
    """lib/trajectory/trajectory_1d.py:76-97."""
    a0, a1, a2 = s0, ds0, dds0 / 2.0
    A = np.array([[3 * T ** 2, 4 * T ** 3], [6 * T, 12 * T ** 2]])
    b = np.array([ds1 - a1 - 2 * a2 * T, dds1 - 2 * a2])
    x = np.linalg.solve(A, b)
    return np.array([a0, a1, a2, x[0], x[1], 0.0])


def poly_eval(c: np.ndarray, t: np.ndarray):
    """Value and three derivatives in the reference's power form (trajectory_1d.py:42-71, 99-130).
    c: [..., 6] (a5 = 0 for a quartic), t: [N] or broadcastable.  Returns 4 arrays [..., N]."""
    a0, a1, a2, a3, a4, a5 = (c[..., k, None] for k in range(6))
    t2, t3, t4, t5 = t ** 2, t ** 3, t ** 4, t ** 5
    p = a0 + a1 * t + a2 * t2 + a3 * t3 + a4 * t4 + a5 * t5
    d1 = a1 + 2 * a2 * t + 3 * a3 * t2 + 4 * a4 * t3 + 5 * a5 * t4
    d2 = 2 * a2 + 6 * a3 * t + 12 * a4 * t2 + 20 * a5 * t3
    d3 = 6 * a3 + 24 * a4 * t + 60 * a5 * t2
    return p, d1, d2, d3


def _seq_sum_sq(x: np.ndarray) -> np.ndarray:
    """`sum(np.power(lst, 2))` (trajectory_planner.py:143): strict left-to-right accumulation."""
    return np.cumsum(x * x, axis=-1)[..., -1]


@dataclass
class Lattice:
    """Everything `generate_range_polynomials` produces, as padded arrays (NaN beyond N)."""
    D: np.ndarray
    T: np.ndarray
    V: np.ndarray
    nsteps_T: np.ndarray          # [n_T]
    nsteps: np.ndarray            # [C]
    row_of: np.ndarray            # [C]
    t: np.ndarray                 # [R, Nmax]   absolute time (t_initial added)
    lon: np.ndarray               # [R, 4, Nmax]  s, ds, dds, ddds
    lat: np.ndarray               # [C, 4, Nmax]  d, dd, ddd, dddd
    clon: np.ndarray              # [R, 6]
    clat: np.ndarray              # [C, 6]
    cd: np.ndarray
    cv: np.ndarray
    ct: np.ndarray
    ctot: np.ndarray
    keep: np.ndarray              # [C] bool; False = candidate dropped (Optimal low-speed, S <= S_thresh)
    low_speed: np.ndarray         # [R] bool
    S: np.ndarray                 # [R]
    feasible: np.ndarray = None   # [C] bool (extension; all True for infinite limits)
    extras: dict = field(default_factory=dict)

    @property
    def n_candidates(self):
        return self.nsteps.shape[0]

    def kept_indices(self):
        return np.nonzero(self.keep)[0]


def generate(prm: OracleParams, p0, s0, t_initial: float, dd: float = 0.0, desired_speed: float = None,
             target_triples: np.ndarray = None, di_interval=None, t_interval=None) -> Lattice:
    """Restates `generate_range_polynomials` (trajectory_planner.py:108-181 and the variant files).

    target_triples: None for velocity keeping, else [n_T][3] = (st, dst, ddst) per horizon as the
    reference computes them on lines 132-136 (V1) / trajectory_planner_optimal.py:130-133.
    """
    variant = prm.variant
    desired_speed = prm.desired_speed if desired_speed is None else desired_speed
    p0v, dp0, ddp0 = (float(v) for v in p0)
    s0v, ds0, dds0 = (float(v) for v in s0)
    follow = target_triples is not None
    di = prm.di_interval() if di_interval is None else di_interval
    ti = prm.t_interval() if t_interval is None else t_interval
    period = prm.sample_period()

    D = np.arange(di[0], di[1], di[2])
    T = np.arange(ti[0], ti[1], ti[2])
    V = long_axis(prm, ds0, follow)
    nV, nT, nD = len(V), len(T), len(D)
    R, C = nV * nT, nV * nT * nD
    tlists = [sample_times(Tj, period) for Tj in T]
    nsteps_T = np.array([len(x) for x in tlists], dtype=np.int64)
    Nmax = int(nsteps_T.max()) if nT else 0

    t = np.full((R, Nmax), np.nan)
    lon = np.full((R, 4, Nmax), np.nan)
    lat = np.full((C, 4, Nmax), np.nan)
    clon = np.zeros((R, 6))
    clat = np.zeros((C, 6))
    cd = np.zeros(C)
    cv = np.zeros(C)
    ct = np.zeros(C)
    ctot = np.zeros(C)
    keep = np.ones(C, dtype=bool)
    low = np.zeros(R, dtype=bool)
    Sarr = np.zeros(R)
    nsteps = np.zeros(C, dtype=np.int64)
    row_of = np.repeat(np.arange(R), nD)

    for iv, v in enumerate(V):
        for iT, Tj in enumerate(T):
            r = iv * nT + iT
            tl = tlists[iT]
            N = len(tl)
            if follow:   # trajectory_planner.py:131-144
                st, dst, ddst = (float(x) for x in target_triples[iT])
                cl = quintic_coeffs(s0v, ds0, dds0, st + v, dst, ddst, Tj)[0]
            else:        # :146-154
                cl = quartic_coeffs(s0v, ds0, dds0, v, 0.0, Tj)
            s, ds, dds, ddds = (x[0] for x in poly_eval(cl[None, :], tl))
            jl = _seq_sum_sq(ddds)
            if follow:
                c_long = prm.kj * jl + prm.kt * Tj + prm.ks * (s[-1] - st) ** 2
            else:
                c_long = prm.kj * jl + prm.kt * Tj + prm.kdots * (ds[-1] - desired_speed) ** 2
            clon[r] = cl
            lon[r, :, :N] = (s, ds, dds, ddds)
            t[r, :N] = tl + t_initial          # :178-179
            if variant == VARIANT_OPTIMAL:     # trajectory_planner_optimal.py:151
                S = s[-1] - s0v
            else:                              # trajectory_planner.py:155
                S = abs(s[-1] - s0v)
            Sarr[r] = S
            sl = slice(r * nD, (r + 1) * nD)
            nsteps[sl] = N
            if follow:
                ct[sl] = c_long
            else:
                cv[sl] = c_long

            # Which lateral parameterisation (trajectory_planner.py:159, variants :156, optimal :155-168)
            if variant == VARIANT_V1:
                use_low, dropped = (ds0 < prm.low_speed_threshold and S > 0), False
            elif variant in (VARIANT_DT, VARIANT_DTOBS):
                use_low, dropped = False, False
            else:
                use_low = ds0 < prm.low_speed_threshold
                dropped = use_low and not (S > prm.S_thresh)
            low[r] = use_low
            if dropped:
                keep[sl] = False
                continue
            if use_low:
                horizon = S
                arg = np.abs(s - s0v)
            else:
                horizon = Tj
                arg = tl
            cq = quintic_coeffs(p0v, dp0, ddp0, D, 0.0, 0.0, horizon)       # [nD, 6]
            dval, d1, d2, d3 = poly_eval(cq, arg)
            jd = _seq_sum_sq(d3)
            c_lat = prm.kj * jd + prm.kt * horizon + prm.kd * (D - dd) ** 2
            clat[sl] = cq
            lat[sl, 0, :N], lat[sl, 1, :N], lat[sl, 2, :N], lat[sl, 3, :N] = dval, d1, d2, d3
            cd[sl] = c_lat
            ctot[sl] = prm.klat * c_lat + prm.klong * c_long

    L = Lattice(D=D, T=T, V=V, nsteps_T=nsteps_T, nsteps=nsteps, row_of=row_of, t=t, lon=lon, lat=lat,
                clon=clon, clat=clat, cd=cd, cv=cv, ct=ct, ctot=ctot, keep=keep, low_speed=low, S=Sarr)
    L.feasible = kinematic_feasible(prm, L)
    return L


def kinematic_feasible(prm: OracleParams, L: Lattice) -> np.ndarray:
    """EXTENSION (no reference code; Werling 2010 Sec. VI): a candidate is kinematically feasible
    iff max_k |ds_k| <= v_max, max_k |dds_k| <= a_max and max_k |ddd_k| <= a_max over its samples.
    With infinite limits every candidate is feasible, which is the reference's behaviour."""
    with np.errstate(invalid="ignore"):
        vmax = np.nanmax(np.abs(L.lon[:, 1, :]), axis=1)
        amax_s = np.nanmax(np.abs(L.lon[:, 2, :]), axis=1)
        amax_d = np.where(L.keep, np.nanmax(np.abs(np.where(np.isnan(L.lat[:, 2, :]), 0.0, L.lat[:, 2, :])), axis=1), 0.0)
    ok_row = (vmax <= prm.v_max) & (amax_s <= prm.a_max)
    return ok_row[L.row_of] & (amax_d <= prm.a_max)


def first_argmin(values: np.ndarray, mask: np.ndarray = None) -> int:
    """`min(paths, key=attrgetter(k))` (trajectory_planner.py:99-106, 212): first minimal element wins.
    Returns -1 when the mask leaves nothing (the reference raises ValueError from `min([])`)."""
    if mask is None:
        return int(np.argmin(values)) if len(values) else -1
    idx = np.nonzero(mask)[0]
    if idx.size == 0:
        return -1
    return int(idx[np.argmin(values[idx])])


def optimal_at_time(time: float, t: np.ndarray, x: np.ndarray, dx: np.ndarray, ddx: np.ndarray, period: float):
    """trajectory_planner.py:183-195 (DTObstacles divides by sampling_t, trajectory_planner_obstacles.py:197).
    t, x, dx, ddx: the winner's unpadded sample arrays."""
    if time <= t[0]:
        k = 0
    elif time >= t[-1]:
        k = -1
    else:
        k = round((time - t[0]) / period)
    return (x[k], dx[k], ddx[k])


# --------------------------------------------------------------------------------------------
# Reference paths (K3)
# --------------------------------------------------------------------------------------------
@dataclass
class CirclePath:
    """lib/trajectory/circle_trajectory.py:11-72 (rounded rectangle, arc-length parameter)."""
    l: float
    h: float
    r: float
    dt: float = 0.1

    def pt(self, t):
        l, h, r = self.l, self.h, self.r
        t = np.array(t, dtype=np.float64, copy=True)
        period = 2 * l + 2 * r * math.pi + 2 * h
        wrap = t >= period
        t[wrap] = t[wrap] % period
        b1 = l + 2 * r * math.pi / 4
        b2 = l + 2 * r * math.pi / 4 + h
        b3 = l + r * math.pi + h
        b4 = 2 * l + r * math.pi + h
        b5 = 2 * l + r * math.pi + h + 2 * r * math.pi / 4
        b6 = 2 * l + r * math.pi + 2 * h + 2 * r * math.pi / 4
        conds = [(t >= 0) & (t < l), (t >= l) & (t < b1), (t >= b1) & (t < b2), (t >= b2) & (t < b3),
                 (t >= b3) & (t < b4), (t >= b4) & (t < b5), (t >= b5) & (t < b6)]
        a1 = (t - l) / r
        a3 = (t - b2) / r
        a5 = (t - b4) / r
        a7 = (t - b6) / r
        xs = [t, l + r * np.sin(a1), np.full_like(t, l + r), l + np.cos(a3) * r,
              l - (t - l - r * math.pi - h), -np.sin(a5) * r, np.full_like(t, -r)]
        ys = [np.zeros_like(t), r - r * np.cos(a1), t - l - 2 * r * math.pi / 4 + r, r + h + np.sin(a3) * r,
              np.full_like(t, 2 * r + h), r + h + r * np.cos(a5), r + h - (t - b5)]
        x = np.select(conds, xs, default=-r * np.cos(a7))
        y = np.select(conds, ys, default=r - r * np.sin(a7))
        return x, y

    def tangent(self, t):
        t = np.asarray(t, dtype=np.float64)
        x1, y1 = self.pt(t + self.dt)
        x0, y0 = self.pt(t)
        return (x1 - x0) / self.dt, (y1 - y0) / self.dt


@dataclass
class SplinePath:
    """lib/trajectory/spline_trajectory.py:9-158 (natural cubic spline over chord length)."""
    wx: list
    wy: list

    def __post_init__(self):
        dx, dy = np.diff(self.wx), np.diff(self.wy)
        ds = [math.sqrt(ix ** 2 + iy ** 2) for ix, iy in zip(dx, dy)]
        s = [0]
        s.extend(np.cumsum(ds))
        self.knots = np.asarray(s, dtype=np.float64)
        self.cx = self._coeffs(self.knots, self.wx)
        self.cy = self._coeffs(self.knots, self.wy)

    @staticmethod
    def _coeffs(x, y):
        n = len(x)
        h = np.diff(x)
        a = np.asarray(y, dtype=np.float64)
        A = np.zeros((n, n))
        A[0, 0] = 1.0
        for i in range(n - 1):
            if i != n - 2:
                A[i + 1, i + 1] = 2.0 * (h[i] + h[i + 1])
            A[i + 1, i] = h[i]
            A[i, i + 1] = h[i]
        A[0, 1] = 0.0
        A[n - 1, n - 2] = 0.0
        A[n - 1, n - 1] = 1.0
        B = np.zeros(n)
        for i in range(n - 2):
            B[i + 1] = 3.0 * (a[i + 2] - a[i + 1]) / h[i + 1] - 3.0 * (a[i + 1] - a[i]) / h[i]
        c = np.linalg.solve(A, B)
        d = np.zeros(n)
        b = np.zeros(n)
        for i in range(n - 1):
            d[i] = (c[i + 1] - c[i]) / (3.0 * h[i])
            b[i] = (a[i + 1] - a[i]) / h[i] - h[i] * (c[i + 1] + 2.0 * c[i]) / 3.0
        return a, b, c, d

    def _eval(self, co, t, deriv):
        a, b, c, d = co
        k = self.knots
        t = np.asarray(t, dtype=np.float64)
        i = np.clip(np.searchsorted(k, t, side="right") - 1, 0, len(k) - 2)
        dx = t - k[i]
        if deriv == 0:
            val = a[i] + b[i] * dx + c[i] * dx ** 2.0 + d[i] * dx ** 3.0
        else:
            val = b[i] + 2.0 * c[i] * dx + 3.0 * d[i] * dx ** 2.0
        val = np.where(t < k[0], k[0], val)       # out of range: the knot abscissa (:47-50, :65-68)
        val = np.where(t >= k[-1], k[-1], val)
        return val

    def pt(self, t):
        return self._eval(self.cx, t, 0), self._eval(self.cy, t, 0)

    def tangent(self, t):
        return self._eval(self.cx, t, 1), self._eval(self.cy, t, 1)


def frenet_to_cartesian(path, s: np.ndarray, d: np.ndarray):
    """lib/tests/test_obstacle_planner.py:18-25, 53-56: (x, y) = pt(s) + (-sin th, cos th) * d with
    th = arctan2 of the path tangent at s.  s and d broadcast against each other."""
    px, py = path.pt(s)
    gx, gy = path.tangent(s)
    th = np.arctan2(gy, gx)
    return px + (-np.sin(th)) * d, py + np.cos(th) * d


def lattice_to_cartesian(path, L: Lattice):
    """x, y [C, Nmax] for every candidate of a lattice (NaN beyond N and for dropped candidates)."""
    s = L.lon[L.row_of, 0, :]
    with np.errstate(invalid="ignore"):
        return frenet_to_cartesian(path, np.where(np.isnan(s), 0.0, s), L.lat[:, 0, :])


# --------------------------------------------------------------------------------------------
# Collision (K4) and sensor gate
# --------------------------------------------------------------------------------------------
def check_collisions(x: np.ndarray, y: np.ndarray, nsteps: np.ndarray, obstacles: np.ndarray,
                     radius_sq: float = ROBOT_RADIUS ** 2):
    """lib/tests/test_moving_obstacle_planner.py:42-52 for every candidate at once.
    obstacles: [O][2] snapshot (reference behaviour) or [K][O][2] time-indexed (EXTENSION, unpinned:
    step k is tested against table k, the last table is reused beyond K).
    Returns (ok [C] bool - True = collision free, first_blocker [C] int - index in list order or -1)."""
    C, Nmax = x.shape
    obstacles = np.asarray(obstacles, dtype=np.float64)
    first = np.full(C, -1, dtype=np.int64)
    if obstacles.size == 0:
        return np.ones(C, dtype=bool), first
    valid = np.arange(Nmax)[None, :] < nsteps[:, None]
    n_obs = obstacles.shape[-2]
    for o in range(n_obs - 1, -1, -1):
        if obstacles.ndim == 2:
            ox, oy = obstacles[o, 0], obstacles[o, 1]
        else:
            kk = np.minimum(np.arange(Nmax), obstacles.shape[0] - 1)
            ox, oy = obstacles[kk, o, 0][None, :], obstacles[kk, o, 1][None, :]
        with np.errstate(invalid="ignore"):
            hit = (((x - ox) ** 2 + (y - oy) ** 2) <= radius_sq) & valid
        first[hit.any(axis=1)] = o
    return first < 0, first


def sensor_gate(robot_pose, obstacles_xy: np.ndarray, rng: float, aperture: float) -> np.ndarray:
    """lib/sensor/proximity_sensor.py:41-75: cone gate, returns a bool mask in list order."""
    p = np.asarray(robot_pose[:2], dtype=np.float64)
    th = robot_pose[2]
    direction = np.array([np.cos(th), np.sin(th)])
    out = np.zeros(len(obstacles_xy), dtype=bool)
    for i, o in enumerate(obstacles_xy):
        diff = np.asarray(o, dtype=np.float64) - p
        cone_dist = np.dot(diff, direction)
        if cone_dist < 0.0 or cone_dist >= rng:
            continue
        cosang = np.dot(direction, diff)
        sinang = abs(direction[0] * diff[1] - direction[1] * diff[0])   # |cross| of two 2-vectors
        if np.abs(np.arctan2(sinang, cosang)) > aperture:
            continue
        out[i] = True
    return out


def curvature_feasible(x: np.ndarray, y: np.ndarray, nsteps: np.ndarray, kappa_max: float) -> np.ndarray:
    """EXTENSION (unpinned): discrete (Menger) curvature of the Cartesian polyline at interior samples,
    kappa = 2 |a x b| / (|a| |b| |a + b|); compared squared and division-free:
    feasible iff 4 (a x b)^2 <= kappa_max^2 |a|^2 |b|^2 |a+b|^2 for every interior k."""
    C, Nmax = x.shape
    ok = np.ones(C, dtype=bool)
    if not np.isfinite(kappa_max):
        return ok
    ax, ay = x[:, 1:-1] - x[:, :-2], y[:, 1:-1] - y[:, :-2]
    bx, by = x[:, 2:] - x[:, 1:-1], y[:, 2:] - y[:, 1:-1]
    cx, cy = x[:, 2:] - x[:, :-2], y[:, 2:] - y[:, :-2]
    cr = ax * by - ay * bx
    lhs = 4.0 * (cr * cr)
    rhs = (kappa_max * kappa_max) * ((ax * ax + ay * ay) * (bx * bx + by * by) * (cx * cx + cy * cy))
    valid = (np.arange(1, Nmax - 1)[None, :] + 1) < nsteps[:, None]
    with np.errstate(invalid="ignore"):
        bad = (lhs > rhs) & valid
    return ~bad.any(axis=1)


# --------------------------------------------------------------------------------------------
# Moving obstacles and targets (host-side input producers)
# --------------------------------------------------------------------------------------------
def obstacle_time_law(t, speed, time_offset):
    """lib/sensor/dynamic_obstacle.py:43-44: `time_offset + t * speed % 50`."""
    return time_offset + (t * speed) % 50


def moving_obstacle_position(path, t, offset, speed, time_offset):
    """lib/sensor/dynamic_obstacle.py:46-56: path(time_law(t)) + n(time_law(t)) * offset."""
    tau = obstacle_time_law(np.asarray(t, dtype=np.float64), speed, time_offset)
    return frenet_to_cartesian(path, tau, offset)


def target_frenet(s0, s1, Ts, DTs=0.1):
    """lib/optimal_frenet_frame/target.py:17-24: leading-vehicle quintic sampled every DTs."""
    c = quintic_coeffs(s0[0], s0[1], s0[2], s1[0], s1[1], s1[2], Ts)[0]
    t = np.arange(0, Ts + DTs, DTs)
    s, ds, dds, ddds = (x[0] for x in poly_eval(c[None, :], t))
    return {"t": t, "s": s, "dot_s": ds, "ddot_s": dds, "dddot_s": ddds}


def follow_target(tg, D0=1.0, DTs=0.1):
    """target.py:29-36."""
    out = dict(tg)
    out["s"] = tg["s"] - D0 + DTs * tg["dot_s"]
    out["dot_s"] = tg["dot_s"] - DTs * tg["ddot_s"]
    if not np.all(tg["ddot_s"] == tg["ddot_s"][0]):
        out["ddot_s"] = tg["ddot_s"] - DTs * tg["dddot_s"]
    return out


def merge_target(tg_a, tg_b):
    """target.py:38-46."""
    out = dict(tg_a)
    out["s"] = 0.5 * (tg_a["s"] + tg_b["s"])
    return out


def optimal_target_triples(tg, T: np.ndarray, t_initial: float, delta_t: float) -> np.ndarray:
    """trajectory_planner_optimal.py:130-133: index the sampled target at t_initial + T."""
    out = np.zeros((len(T), 3))
    for j, Tj in enumerate(T):
        k = round((t_initial + Tj - tg["t"][0]) / delta_t)
        out[j] = (tg["s"][k], tg["dot_s"][k], tg["ddot_s"][k])
    return out


# --------------------------------------------------------------------------------------------
# Duckietown lane runs: parabola reference path, off-road filter
# --------------------------------------------------------------------------------------------
AGENT_SAFETY_RAD = (max(0.18, 0.13 + 0.02) / 2) * 1.8      # lib/video/constants.py:27-33


@dataclass
class ParabolaPath:
    """The lane-centre path the mapper fits in the robot frame (lib/mapper/mapper_obstacles.py:169-186):
    pt(x) = (x, a x^2 + b x + c), parameterised by x.  `tangent` is the chord over ds = 1/30 that the Duckietown
    driver's compute_ortogonal_vect uses (lib/tests/test_mapper_planner_obstacles.py:137-142), NOT divided by ds."""
    a: float
    b: float
    c: float
    ds: float = 1 / 30

    def pt(self, x):
        x = np.asarray(x, dtype=np.float64)
        return x, self.a * x ** 2 + self.b * x + self.c

    def tangent(self, x):
        x = np.asarray(x, dtype=np.float64)
        x1 = x + self.ds
        return x1 - x, (self.a * x1 ** 2 + self.b * x1 + self.c) - (self.a * x ** 2 + self.b * x + self.c)


def vertex_form(fit):
    """get_vertex_and_focus_distance (test_mapper_planner_obstacles.py:105-109): (h, k, a) of a x^2 + b x + c."""
    a, b, c = fit
    return -b / (2 * a), -(b ** 2 - 4 * a * c) / (4 * a), a


def offroad(x: np.ndarray, y: np.ndarray, nsteps: np.ndarray, fit, side: str) -> np.ndarray:
    """check_offroad (test_mapper_planner_obstacles.py:62-70) for every candidate: True = leaves the road on `side`."""
    h, k, a = vertex_form(fit)
    valid = np.arange(x.shape[1])[None, :] < nsteps[:, None]
    with np.errstate(invalid="ignore"):
        dist = (y - k) - a * (x - h) ** 2
        bad = (dist < 0) if side == "right" else (dist > 0)
    return (bad & valid).any(axis=1)
