"""Shared harness of the GPU parity tests: drives the CUDA path through the engine / C ABI."""
import numpy as np

import otg_b200  # noqa: F401
from otg_b200 import _native as nat
from otg_b200 import lattice as lat
from otg_b200.engine import FrenetEngine, LatticeGeometry, make_params

from oracle import frenet_oracle as fo


class Attr:
    def __init__(self, d):
        self.__dict__.update(d)


def planner_attrs(prm: fo.OracleParams) -> Attr:
    """OracleParams -> object with the reference's attribute names."""
    return Attr(dict(kj=prm.kj, kt=prm.kt, kd=prm.kd, ks=prm.ks, kdots=prm.kdots, klat=prm.klat, klong=prm.klong,
                     low_speed_threshold=prm.low_speed_threshold, S_thresh=prm.S_thresh, delta_t=prm.delta_t,
                     sampling_t=prm.sampling_t, dt=prm.dt, min_t=prm.min_t, max_t=prm.max_t,
                     max_road_width=prm.max_road_width, d_road_width=prm.d_road_width, d_d_s=prm.d_d_s,
                     num_sample=prm.num_sample, v_max=prm.v_max, a_max=prm.a_max, kappa_max=prm.kappa_max,
                     desired_speed=prm.desired_speed))


def run_batch(prm: fo.OracleParams, p0, s0, t0, dd=None, desired=None, targets=None, trajectory=None, obstacles=None,
              active=None, use_mask=0, di_interval=None, t_interval=None, gather_sel=nat.SEL_CTOT, curvature=False,
              engine=None, no_fuse=False, projection=None, radius_sq=0.7 ** 2, one_launch=False, repeat=1):
    """Runs B scenarios through K1..K5 (+K3/K4 when a path / obstacles are given).  Returns the engine."""
    a = planner_attrs(prm)
    p0 = np.atleast_2d(np.asarray(p0, dtype=np.float64))
    s0 = np.atleast_2d(np.asarray(s0, dtype=np.float64))
    B = p0.shape[0]
    t0 = np.broadcast_to(np.asarray(t0, dtype=np.float64), (B,))
    dd = np.zeros(B) if dd is None else np.broadcast_to(np.asarray(dd, dtype=np.float64), (B,))
    desired = np.full(B, prm.desired_speed) if desired is None else np.broadcast_to(np.asarray(desired, dtype=np.float64), (B,))
    follow = targets is not None
    di = lat.di_interval(a, prm.variant) if di_interval is None else di_interval
    ti = lat.t_interval(a, prm.variant) if t_interval is None else t_interval
    geom = LatticeGeometry.from_intervals(di, ti, lat.sample_period(a, prm.variant))
    axes = [lat.long_axis(a, s0[b, 1], follow) for b in range(B)]
    n_v_cap = max(len(x) for x in axes)
    n_obs = 0 if obstacles is None else np.asarray(obstacles).shape[1]
    eng = engine or FrenetEngine(B)
    eng.configure(geom, n_v_cap, n_obs)
    eng.set_relative_s(projection is not None)       # Duckietown lane runs: path evaluated at projection + (s - s0[0])
    if projection is not None:
        eng.path_origin[:, 0], eng.path_origin[:, 1] = np.broadcast_to(np.asarray(projection, dtype=np.float64), (B,)), s0[:, 0]
    eng.state[:, :3], eng.state[:, 3:] = p0, s0
    eng.scal[:, 0], eng.scal[:, 1], eng.scal[:, 2] = dd, desired, t0
    eng.axis_v[:] = 0
    for b, x in enumerate(axes):
        eng.axis_v[b, :len(x)] = x
        eng.n_v[b] = len(x)
    if follow:
        eng.target[:] = np.asarray(targets, dtype=np.float64).reshape(B, geom.n_t, 3)
    stages = nat.STAGE_K1 | nat.STAGE_K2 | nat.STAGE_K5 | nat.STAGE_GATHER
    path = None
    if trajectory is not None:
        path = eng.path_handle(trajectory)
        stages |= nat.STAGE_K3
        if curvature:
            stages |= nat.STAGE_CURVATURE
            use_mask |= nat.MASK_CURVATURE
    if obstacles is not None:
        eng.obs_xy[:] = np.asarray(obstacles, dtype=np.float64)
        eng.obs_active[:] = 1 if active is None else np.asarray(active, dtype=np.uint8)
        stages |= nat.STAGE_K4
        use_mask |= nat.MASK_COLLISION
    if no_fuse:
        stages |= nat.STAGE_NO_FUSE
    if one_launch:
        stages |= nat.STAGE_ONE_LAUNCH
    P = make_params(a, lat.VARIANTS[prm.variant]["rule"], lat.sample_period(a, prm.variant), follow, radius_sq=radius_sq)
    for _ in range(repeat):
        eng.run(P, path, stages, use_mask, gather_sel)
    return eng


def compact(fields: dict) -> dict:
    """Drop candidates the reference would not have appended (Optimal variant) -> golden list order."""
    v = fields["valid"]
    return {k: (a[v] if isinstance(a, np.ndarray) and a.shape[:1] == v.shape else a) for k, a in fields.items()}
