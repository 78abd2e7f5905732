"""Planner side of the Duckietown lane runs.

Mirrors the planner-facing functions of lib/tests/test_mapper_planner_obstacles.py:25-142 - the projection-relative
Frenet -> Cartesian transform, the off-road filter against the two lane-boundary parabolas, the two-point-set collision
filter with the Duckiebot safety radius, and the cheapest survivor - on the lattice that the planner's replan left in HBM.
Perception (camera image -> lane fits, obstacle boxes, `mapper.cam2rob`) and the gym simulator are outside the hot path:
the functions take what `mapper.process` returns, exactly like the reference's.

    paths = frenet_to_glob_planner(planner, trajectory, planner.paths, est_pt)   # K3 at est_pt + (s - s0[0])
    paths = check_paths(paths, obstacles, robot_p, mapper, rw, lw)               # off-road + K4 + K5 + winner gather
    planner.opt_path_tot = min(paths, key=attrgetter('ctot'))                    # already known from the K5 record
"""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from ..engine import make_path_handle
from ..trajectory.paths import ParabolaTrajectory2D, PATH_KIND_PARABOLA
from .common import _is_live, _engine_for_lists

# lib/video/constants.py:27-33 (DuckietownParameters)
ROBOT_LENGTH, ROBOT_WIDTH, SAFETY_RAD_MULT = 0.18, 0.13 + 0.02, 1.8
AGENT_SAFETY_RAD = (max(ROBOT_LENGTH, ROBOT_WIDTH) / 2) * SAFETY_RAD_MULT


def lane_trajectory(trajectory) -> ParabolaTrajectory2D:
    """The lane-centre parabola behind a path object.  Ours is returned as is; the reference's bare `Trajectory()` with
    lambdas (`build_trajectory`, lib/mapper/mapper_obstacles.py:169-186) is recovered from three evaluations: c = y(0),
    b = y'(0), a = y''/2 - the coefficients the lambdas close over, bit for bit."""
    if isinstance(trajectory, ParabolaTrajectory2D):
        return trajectory
    if hasattr(trajectory, "compute_second_derivative"):
        c = float(trajectory.compute_pt(0.0)[1])
        b = float(trajectory.compute_first_derivative(0.0)[1])
        a = float(trajectory.compute_second_derivative(0.0)[1]) / 2
        return ParabolaTrajectory2D(a, b, c)
    raise TypeError(f"{type(trajectory).__name__} is not a lane-centre parabola")


def compute_ortogonal_vect(trajectory, s):
    """Unit normal from the chord over ds = 1/30 (test_mapper_planner_obstacles.py:137-142): one point, host side."""
    ds = 1 / 30
    s1 = s + ds
    t_grad = trajectory.compute_pt(s1) - trajectory.compute_pt(s)
    t_r = np.arctan2(t_grad[1], t_grad[0])
    return np.array([-np.sin(t_r), np.cos(t_r)])


def _lane_handle(eng, trajectory):
    lane = lane_trajectory(trajectory)
    return make_path_handle(lane, eng.device)     # a parabola descriptor owns no device memory: nothing to cache


def frenet_to_glob_planner(planner, trajectory, frenet_paths, projection):
    """x, y of every candidate with the path evaluated at projection + (s - planner.s0[0]) (:25-37): K3."""
    if _is_live(frenet_paths):
        eng = frenet_paths._engine
        eng.set_relative_s(True)
        eng.path_origin[0] = (float(projection), float(planner.s0[0]))
        eng.upload_region("path_origin")
        eng.launch(planner._params, _lane_handle(eng, trajectory), nat.STAGE_K3)
        eng.xy_path = None          # (a lane transform: drivers.frenet_to_glob must not take these x, y for a path's)
        return frenet_paths
    paths = list(frenet_paths)
    if paths:
        eng = _engine_for_lists()
        s0 = planner.s0[0]
        s = np.concatenate([[projection + (si - s0) for si in p.s] for p in paths]).astype(np.float64)
        d = np.concatenate([np.asarray(p.d, dtype=np.float64) for p in paths])
        x, y = eng.points_to_cartesian(_lane_handle(eng, trajectory), s, d)
        at = 0
        for p in paths:
            n = len(p.s)
            p.x, p.y = x[at:at + n].tolist(), y[at:at + n].tolist()      # the reference resets x, y here (:27-28)
            at += n
    return frenet_paths


def frenet_to_glob(trajectory, planner, projection):
    """[N][2] Cartesian samples of planner.opt_path_tot (:39-49)."""
    f = planner.opt_path_tot
    eng = planner._engine
    s0 = planner.s0[0]
    s = np.array([projection + (si - s0) for si in f.s], dtype=np.float64)
    x, y = eng.points_to_cartesian(_lane_handle(eng, trajectory), s, np.asarray(f.d, dtype=np.float64))
    return np.stack([x, y], axis=1)


def get_vertex_and_focus_distance(fit):
    """(h, k, a) of y = a x^2 + b x + c (:105-109)."""
    a, b, c = fit
    h = -b / (2 * a)
    k = -(b ** 2 - 4 * a * c) / (4 * a)
    return h, k, a


def obstacles_coordinates(obstacles, mapper):
    """Five robot-frame point sets of the detected obstacles (:72-103): end_point, center, end_right, end_lat, end_top."""
    keys = ("end_point", "center", "end_right", "end_lat", "end_top")
    if len(obstacles) == 0:
        return [], [], [], [], []
    return tuple(mapper.cam2rob(np.array([ob[k] for ob in obstacles])) for k in keys)


def check_collisions(path, obstacles):
    """True = `path` keeps AGENT_SAFETY_RAD from every obstacle point (:51-60).  One path: the polyline kernel."""
    if len(obstacles) == 0:
        return True
    free, _ = _engine_for_lists().polyline_collision([path.x], [path.y], np.asarray(obstacles, dtype=np.float64).reshape(-1, 2),
                                                     AGENT_SAFETY_RAD ** 2)
    return bool(free[0])


def check_offroad(path, params, side):
    """True = one sample of `path` lies beyond the boundary parabola (h, k, a) on `side` (:62-70).  Single-path helper kept
    for API parity (a handful of host flops); `check_paths` runs the same test for the whole lattice on the device."""
    h, k, a = params
    x, y = np.asarray(path.x, dtype=np.float64), np.asarray(path.y, dtype=np.float64)
    distance = (y - k) - a * (x - h) ** 2
    return bool((distance < 0).any() if side == "right" else (distance > 0).any())


def _given(fit) -> bool:
    return bool(np.array(fit != None).all())      # noqa: E711 - the reference's own test (:117, :121)


def check_paths(frenet_paths, obstacles, rpose, mapper, rwfit, lwfit):
    """Candidates that stay on the road and clear of the obstacles' right-end and lateral-end points (:112-135)."""
    obs_r, obs_lat = [], []
    if len(obstacles):
        _, _, obs_r, obs_lat, _ = obstacles_coordinates(obstacles, mapper)
    rw = rwfit if _given(rwfit) else None
    lw = lwfit if _given(lwfit) else None
    if _is_live(frenet_paths):
        return _check_lattice(frenet_paths, np.asarray(obs_r, dtype=np.float64).reshape(-1, 2),
                              np.asarray(obs_lat, dtype=np.float64).reshape(-1, 2), rw, lw)
    kept = []
    for p in frenet_paths:
        if rw is not None and check_offroad(p, get_vertex_and_focus_distance(rw), "right"):
            continue
        if lw is not None and check_offroad(p, get_vertex_and_focus_distance(lw), "left"):
            continue
        if len(obstacles) and not (check_collisions(p, obs_r) and check_collisions(p, obs_lat)):
            continue
        kept.append(p)
    return kept


def _check_lattice(frenet_paths, obs_r, obs_lat, rw, lw):
    eng, planner = frenet_paths._engine, frenet_paths._planner
    if not eng.has_xy:
        raise RuntimeError("check_paths: call frenet_to_glob_planner first (the candidates have no x, y yet)")
    # a path survives iff it clears every right-end point AND every lateral-end point: one K4 pass over the union
    pts = np.concatenate([obs_r, obs_lat])
    n = len(pts)
    if n > eng.n_obs:
        raise ValueError(f"{n} obstacle points exceed the planner's staging capacity {eng.n_obs} (constructor kwarg _n_obs_cap)")
    mask, stages = 0, nat.STAGE_K5 | nat.STAGE_GATHER
    if rw is not None or lw is not None:
        eng.offroad(rw, lw)
        mask |= nat.MASK_ROAD
    if n:
        eng.obs_xy[0, :n] = pts
        eng.obs_active[0, :n] = 1
        eng.upload_region("obs_xy", "obs_active")
        stages |= nat.STAGE_K4
        mask |= nat.MASK_COLLISION
    P = nat.FrenetParams.from_buffer_copy(planner._params)
    P.radius_sq = AGENT_SAFETY_RAD ** 2
    eng.launch(P, None, stages, mask, nat.SEL_MASKED, n_obs=n)
    eng.download(sync=True)
    C_ = frenet_paths._n_all
    winner = planner._frenet_from_winner("argmin_masked")

    def fetch_mask():
        frenet_paths._check_live()
        keep = np.ones(C_, dtype=bool)
        if mask & nat.MASK_ROAD:
            keep &= eng.road_ok[0, :C_].cpu().numpy() != 0
        if mask & nat.MASK_COLLISION:
            keep &= eng.coll_free[0, :C_].cpu().numpy() != 0
        return keep

    return frenet_paths.masked_view(fetch_mask, int(eng.result[0]["n_survivors"]), {"ctot": winner})


def lane_step(planner, trajectory, projection, time, obstacles, mapper, rwfit, lwfit, rpose=(0.07, 0.0, 0.0)):
    """The planner part of one camera frame of the reference's loop (:222-238): replan, transform, filter, pick.
    Returns (pos_s, pos_d, surviving paths, winner xy or None when every candidate was filtered out)."""
    pos_s, pos_d = planner.replanner(time=time)
    paths = frenet_to_glob_planner(planner, trajectory, planner.paths, projection)
    paths = check_paths(paths, obstacles, np.asarray(rpose), mapper, rwfit, lwfit)
    if len(obstacles):
        if len(paths) == 0:
            return pos_s, pos_d, paths, None
        planner.opt_path_tot = paths.argmin("ctot") if hasattr(paths, "argmin") else min(paths, key=lambda f: f.ctot)
    return pos_s, pos_d, paths, frenet_to_glob(trajectory, planner, projection)
