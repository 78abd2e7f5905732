"""Generate the golden fixtures by running the UNMODIFIED reference (read-only at /root/reference).

Run in the build container only:   python tests/golden/make_golden.py [--only NAME ...]

The reference is pure Python and cannot travel to the GPU box, and it ships no golden vectors
of its own, so its outputs on fixed, seeded inputs are frozen here as .npz files.
Every array below comes straight out of reference objects (planner.paths, check_paths, sensor.sense ...);
nothing from this repository's oracle or CUDA path is involved.  The fixtures are what
`tests/test_oracle_golden.py` pins the oracle against and what the `-m gpu` tests compare the CUDA
path with.  Regenerating needs NumPy only (2.3.5 was used; the reference pins no NumPy version).
"""
import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refimport  # noqa: E402

_refimport.install()

from lib.planner import (TrajectoryPlannerV1, TrajectoryPlannerParams, TrajectoryPlannerV1DT,  # noqa: E402
                         TrajectoryPlannerParamsDT, TrajectoryPlannerV1DTObstacles,
                         TrajectoryPlannerParamsDTObstacles, TrajectoryPlannerOptimal,
                         TrajectoryPlannerParamsOptimal)
from lib.trajectory import CircleTrajectory2D, SplineTrajectory2D  # noqa: E402
from lib.transform import FrenetGNTransformOld  # noqa: E402
from lib.controller import FrenetIOLController  # noqa: E402
from lib.platform import Unicycle  # noqa: E402
from lib.sensor import StaticObstacle, MovingObstacle, ProximitySensor, ObstacleTrajectory  # noqa: E402
from lib.optimal_frenet_frame import Target, OptimalParameters as op  # noqa: E402
import lib.tests.test_obstacle_planner as ref_static  # noqa: E402
import lib.tests.test_moving_obstacle_planner as ref_moving  # noqa: E402
from operator import attrgetter  # noqa: E402

FIELDS = ("t", "d", "dot_d", "ddot_d", "dddot_d", "s", "dot_s", "ddot_s", "dddot_s")
SPLINE_WX = [0, 2.5, 5, 7.5, -3, 2.7]
SPLINE_WY = [0.7, -6, 5, -9.5, 0, 5]


def pack_paths(paths, with_xy=False, nmax=None):
    """Reference Frenet objects -> padded arrays [C][Nmax] (NaN beyond each path's length)."""
    C = len(paths)
    n = np.array([len(f.t) for f in paths], dtype=np.int32)
    nmax = int(n.max()) if nmax is None else nmax
    out = {"nsteps": n}
    names = FIELDS + (("x", "y") if with_xy else ())
    for name in names:
        arr = np.full((C, nmax), np.nan)
        for i, f in enumerate(paths):
            v = getattr(f, name)
            arr[i, :len(v)] = v
        out[name] = arr
    for k in ("cd", "cv", "ct", "ctot"):
        out[k] = np.array([getattr(f, k) for f in paths], dtype=np.float64)
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}.npz  {os.path.getsize(path) / 1e6:.2f} MB")


# ------------------------------------------------------------------------------------------------
def gen_kat_variants():
    """Default-parameter `initialize` for the four planner classes + two off-default states."""
    cases = [
        ("v1", TrajectoryPlannerV1, TrajectoryPlannerParams, dict(t0=0.1, p0=(0, 0, 0), s0=(0, 0, 0))),
        ("dt", TrajectoryPlannerV1DT, TrajectoryPlannerParamsDT, dict(t0=0.1, p0=(0, 0, 0), s0=(0, 0, 0))),
        ("dt_obstacles", TrajectoryPlannerV1DTObstacles, TrajectoryPlannerParamsDTObstacles,
         dict(t0=0.1, p0=(0, 0, 0), s0=(0, 0, 0))),
        ("optimal", TrajectoryPlannerOptimal, TrajectoryPlannerParamsOptimal,
         dict(t0=0, p0=(3, .3, 0), s0=(0, 2, -.5))),
        ("v1", TrajectoryPlannerV1, TrajectoryPlannerParams,
         dict(t0=0.3, p0=(0.4, 0.1, -0.2), s0=(3, 1.2, 0.3), dd=0.5)),          # low-speed branch, dd != 0
        ("v1", TrajectoryPlannerV1, TrajectoryPlannerParams,
         dict(t0=2.0, p0=(-0.7, 0.3, 0.1), s0=(11.0, 2.6, -0.4), dd=-1.0)),     # high-speed branch
        ("optimal", TrajectoryPlannerOptimal, TrajectoryPlannerParamsOptimal,
         dict(t0=0, p0=(3, .3, 0), s0=(0, 1.0, -.5))),                           # drops S <= S_thresh rows
        ("dt", TrajectoryPlannerV1DT, TrajectoryPlannerParamsDT,
         dict(t0=1.7, p0=(0.05, -0.02, 0.01), s0=(2.0, 0.45, 0.05), dd=0.1)),
    ]
    out = {"n_cases": np.array(len(cases))}
    for i, (variant, cls, pcls, init) in enumerate(cases):
        pl = cls(pcls())
        pl.initialize(**init)
        pk = pack_paths(pl.paths)
        for k, v in pk.items():
            out[f"c{i}_{k}"] = v
        out[f"c{i}_variant"] = np.array(variant)
        out[f"c{i}_t0"] = np.array(float(init["t0"]))
        out[f"c{i}_p0"] = np.array(init["p0"], dtype=np.float64)
        out[f"c{i}_s0"] = np.array(init["s0"], dtype=np.float64)
        out[f"c{i}_dd"] = np.array(float(init.get("dd", 0)))
        out[f"c{i}_di_interval"] = np.array(pl.di_interval, dtype=np.float64)
        out[f"c{i}_t_interval"] = np.array(pl.t_interval, dtype=np.float64)
        out[f"c{i}_dsi_interval"] = np.array(pl.dsi_interval, dtype=np.float64)
        for key in ("cd", "cv", "ctot"):
            out[f"c{i}_argmin_{key}"] = np.array(pl.paths.index(min(pl.paths, key=attrgetter(key))))
        print(f"  kat {i} {variant}: {len(pl.paths)} candidates, {int(pk['nsteps'].sum())} cand-ts, "
              f"argmin {int(out[f'c{i}_argmin_ctot'])}")
    save("kat_variants", **out)


# ------------------------------------------------------------------------------------------------
def gen_paths():
    rng = np.random.default_rng(7)
    circ = CircleTrajectory2D(8, 5, 2)
    spl = SplineTrajectory2D(SPLINE_WX, SPLINE_WY)
    period = 2 * 8 + 2 * 2 * np.pi + 2 * 5
    s_c = np.concatenate([rng.uniform(-3, 120, 3000), np.array([0.0, 8.0, period, period - 0.05, -0.05, 7.95]),
                          np.array([8 + np.pi, 13 + np.pi, 13 + 2 * np.pi, 21 + 2 * np.pi])])
    s_s = np.concatenate([rng.uniform(-2, 60, 3000), np.array(spl.s), np.array([0.0, -1e-9, spl.s[-1] - 1e-9])])
    d_c = rng.normal(0, 2, s_c.shape)
    d_s = rng.normal(0, 2, s_s.shape)

    def ev(tr, ss, dd):
        pt = np.array([tr.compute_pt(s) for s in ss])
        g = np.array([tr.compute_first_derivative(s) for s in ss])
        nv = np.array([ref_static.compute_ortogonal_vect(tr, s) for s in ss])
        xy = np.array([tr.compute_pt(s) + ref_static.compute_ortogonal_vect(tr, s) * d for s, d in zip(ss, dd)])
        return pt, g, nv, xy

    pc, gc, nc, xc = ev(circ, s_c, d_c)
    ps, gs, ns, xs = ev(spl, s_s, d_s)
    save("paths", circle_lhr=np.array([8., 5., 2.]), circle_s=s_c, circle_d=d_c, circle_pt=pc, circle_grad=gc,
         circle_normal=nc, circle_xy=xc, spline_wx=np.array(SPLINE_WX, dtype=np.float64),
         spline_wy=np.array(SPLINE_WY, dtype=np.float64), spline_s=s_s, spline_d=d_s, spline_pt=ps, spline_grad=gs,
         spline_normal=ns, spline_xy=xs, spline_knots=np.array(spl.s, dtype=np.float64),
         spline_ax=np.array(spl.sx.a), spline_bx=np.array(spl.sx.b), spline_cx=np.array(spl.sx.c),
         spline_dx=np.array(spl.sx.d), spline_ay=np.array(spl.sy.a), spline_by=np.array(spl.sy.b),
         spline_cy=np.array(spl.sy.c), spline_dy=np.array(spl.sy.d))


# ------------------------------------------------------------------------------------------------
def _sim_objects():
    """The circle configuration (config/config_circle2D.json, decoded by hand - SURVEY.md section 2 row 22)."""
    t_vect = np.arange(0.1, 50, 0.1)
    robot = Unicycle()
    robot.set_initial_pose(np.array([0.0, 0.0, 0.0]))
    trajectory = CircleTrajectory2D(8, 5, 2)
    transformer = FrenetGNTransformOld(0.5, 10)
    controller = FrenetIOLController(.5, 0.0, 5.0, 0.0, 0.0)
    planner = TrajectoryPlannerV1(TrajectoryPlannerParams())
    return t_vect, robot, trajectory, transformer, controller, planner


def _seeded_obstacle_positions(t_vect, trajectory):
    """lib/tests/test_obstacle_planner.py:142-155 (identical in the moving-obstacle driver)."""
    np.random.seed(5)
    t_samples = np.random.choice(t_vect, 10)
    xs, ys = [], []
    for t in t_samples:
        pt = trajectory.compute_pt(t)
        xs.append(pt[0])
        ys.append(pt[1])
    pos = np.array([xs, ys])
    pos[1, :] += np.random.normal(0, 2., pos.shape[1])
    return pos


def gen_sim(kind, full_steps=(0, 1, 2, 60, 150, 300, 498), cost_every=5, cmax=384):
    """Closed-loop run with the reference driver logic (lib/tests/test_planner.py:19-78,
    test_obstacle_planner.py:58-116, test_moving_obstacle_planner.py:60-124), recording every replan."""
    t_vect, robot, trajectory, transformer, controller, planner = _sim_objects()
    n = t_vect.shape[0]
    sensor = None
    mod = None
    extra = {}
    if kind != "full":
        pos = _seeded_obstacle_positions(t_vect, trajectory)
        if kind == "static":
            obstacle_lst = [StaticObstacle(pos[:, i]) for i in range(pos.shape[1])]
            mod = ref_static
        else:
            obstacle_lst = [MovingObstacle(pos[:, i], trajectory) for i in range(pos.shape[1])]
            mod = ref_moving
            extra["obs_offset"] = np.array([o.trajectory.offset for o in obstacle_lst], dtype=np.float64)
            extra["obs_speed"] = np.array([o.trajectory.speed for o in obstacle_lst], dtype=np.float64)
            extra["obs_time_offset"] = np.array([o.trajectory.time_offset for o in obstacle_lst], dtype=np.float64)
        extra["obs_start"] = pos.T.copy()
        sensor = ProximitySensor(3.5, np.pi / 6)
        sensor.attach(robot)
        sensor.set_obstacle(obstacle_lst)

    rec = dict(time=np.zeros(n), p0=np.zeros((n, 3)), s0=np.zeros((n, 3)), ret_s=np.zeros((n, 3)),
               ret_p=np.zeros((n, 3)), dsi=np.zeros((n, 3)), ncand=np.zeros(n, dtype=np.int32),
               argmin=np.zeros(n, dtype=np.int32), ctot_min=np.zeros(n), robot_pose=np.zeros((n, 3)),
               robot_pose_in=np.zeros((n, 3)))
    cost_steps = list(range(0, n, cost_every))
    rec["cost_steps"] = np.array(cost_steps, dtype=np.int32)
    for k in ("cd", "cv", "ctot"):
        rec[k] = np.full((len(cost_steps), cmax), np.nan)
    if sensor is not None:
        rec.update(n_sensed=np.zeros(n, dtype=np.int32), obs_xy=np.full((n, 10, 2), np.nan),
                   obs_all_xy=np.zeros((n, 10, 2)), obs_idx=np.full((n, 10), -1, dtype=np.int32),
                   keep=np.zeros((n, cmax), dtype=np.uint8), first_blocker=np.full((n, cmax), -1, dtype=np.int8),
                   argmin_masked=np.zeros(n, dtype=np.int32), follow_fired=np.zeros(n, dtype=np.uint8))
    full = {}

    robot_p = robot.p
    robot_dp = np.zeros(3)
    transformer.estimatePosition(trajectory, robot_p)
    robot_fpose = transformer.transform(robot_p)
    s0 = (robot_fpose[0], 0.0, 0.0)
    d0 = (robot_fpose[1], 0.0, 0.0)
    planner.initialize(t0=t_vect[0], p0=d0, s0=s0)
    rec["init_p0"] = np.array(d0, dtype=np.float64)
    rec["init_s0"] = np.array(s0, dtype=np.float64)
    pk = pack_paths(planner.paths)
    for k, v in pk.items():
        full[f"init_{k}"] = v
    rec["init_argmin"] = np.array(planner.paths.index(planner.opt_path_tot))

    for i in range(n):
        est_pt = transformer.estimatePosition(trajectory, robot_p)
        robot_fpose = transformer.transform(robot_p)
        rec["robot_pose_in"][i] = robot_p
        pos_s, pos_d = planner.replanner(t_vect[i])
        all_paths = planner.paths
        rec["time"][i] = t_vect[i]
        rec["p0"][i] = planner.p0
        rec["s0"][i] = planner.s0
        rec["ret_s"][i] = pos_s
        rec["ret_p"][i] = pos_d
        rec["dsi"][i] = planner.dsi_interval
        rec["ncand"][i] = len(all_paths)
        rec["argmin"][i] = all_paths.index(planner.opt_path_tot)
        rec["ctot_min"][i] = planner.opt_path_tot.ctot
        if i in cost_steps:
            j = cost_steps.index(i)
            for k in ("cd", "cv", "ctot"):
                rec[k][j, :len(all_paths)] = [getattr(f, k) for f in all_paths]
        if sensor is not None:
            paths_planner = mod.frenet_to_glob(planner, trajectory, all_paths)
            if kind == "moving":
                sensor.step_obstacle(t_vect[i])
            rec["obs_all_xy"][i] = np.array([o.position for o in sensor.obstacles])
            sensed = sensor.sense(robot_p)
            rec["n_sensed"][i] = len(sensed)
            for j, o in enumerate(sensed):
                rec["obs_xy"][i, j] = o.position
                rec["obs_idx"][i, j] = [id(x) for x in sensor.obstacles].index(id(o))
            if kind == "static":
                planner.paths = mod.check_paths(paths_planner, sensor, robot_p)
            else:
                planner.paths, leading = mod.check_paths(paths_planner, sensor, robot_p)
                blocked = [c for c, f in enumerate(all_paths) if id(f) not in {id(g) for g in planner.paths}]
                if sensed:
                    assert len(blocked) == len(leading)
                    for c, ob in zip(blocked, leading):
                        rec["first_blocker"][i, c] = [id(x) for x in sensed].index(id(ob))
                if len(planner.paths) == 0:
                    rec["follow_fired"][i] = 1
                    s_target = leading[0].trajectory
                    s_target.set_transformer(transformer)
                    pos_s, pos_d = planner.replanner(t_vect[i], s_target=s_target)
                    paths_planner = mod.frenet_to_glob(planner, trajectory, planner.paths)
            kept_ids = {id(f) for f in planner.paths}
            rec["keep"][i, :len(all_paths)] = [1 if id(f) in kept_ids else 0 for f in all_paths]
            planner.opt_path_tot = min(planner.paths, key=attrgetter('ctot'))
            rec["argmin_masked"][i] = all_paths.index(planner.opt_path_tot) if not rec["follow_fired"][i] else -1
        if i in full_steps:
            pk = pack_paths(all_paths, with_xy=sensor is not None)
            for k, v in pk.items():
                full[f"s{i}_{k}"] = v
        ts, td = pos_s[0], pos_d[0]
        error = np.array([0, td]) - robot_fpose[:2]
        derror = np.array([pos_s[1], pos_d[1]])
        curvature = trajectory.compute_curvature(est_pt)
        u = controller.compute(robot_fpose, error, derror, curvature)
        robot_p, robot_dp = robot.step(u, 0.1)
        rec["robot_pose"][i] = robot_p
    rec["full_steps"] = np.array(full_steps, dtype=np.int32)
    print(f"  sim {kind}: final pose {rec['robot_pose'][-1]}")
    if sensor is not None:
        print(f"    sensed in {(rec['n_sensed'] > 0).sum()} steps, filtered {int((rec['keep'] == 0).sum() - (cmax * n - rec['ncand'].sum()))}"
              f", argmin changed {(rec['argmin'] != rec['argmin_masked']).sum()}, follow {rec['follow_fired'].sum()}")
    save(f"sim_planner_{kind}", **rec, **full, **extra)


# ------------------------------------------------------------------------------------------------
def _dense_planner():
    return TrajectoryPlannerV1(TrajectoryPlannerParams(d_road_width=0.1, dt=0.1, num_sample=3, delta_t=0.02))


def _dense_reference(args):
    """One dense-lattice replan + Frenet->Cartesian + collision check with the reference code."""
    p0, s0, t0, dd, obs_xy, wx, wy, sample_stride = args
    spl = SplineTrajectory2D(wx, wy)
    pl = _dense_planner()
    pl.initialize(t0=t0, p0=tuple(p0), s0=tuple(s0), dd=dd)
    paths = pl.paths
    ref_moving.frenet_to_glob(pl, spl, paths)
    obstacles = [StaticObstacle(np.array(o)) for o in obs_xy]
    keep = np.zeros(len(paths), dtype=np.uint8)
    first = np.full(len(paths), -1, dtype=np.int16)
    for c, f in enumerate(paths):
        ok, ob = ref_moving.check_collisions(f, obstacles)
        keep[c] = 1 if ok else 0
        if not ok:
            first[c] = [id(x) for x in obstacles].index(id(ob))
    out = dict(nsteps=np.array([len(f.t) for f in paths], dtype=np.int32),
               cd=np.array([f.cd for f in paths]), cv=np.array([f.cv for f in paths]),
               ctot=np.array([f.ctot for f in paths]), keep=keep, first_blocker=first,
               argmin=np.array(paths.index(pl.opt_path_tot)),
               dsi_interval=np.array(pl.dsi_interval, dtype=np.float64))
    survivors = [f for f, k in zip(paths, keep) if k]
    out["argmin_masked"] = np.array(paths.index(min(survivors, key=attrgetter('ctot'))) if survivors else -1)
    # per-candidate checksums (plain sums over the samples) of every field + a strided sample of full arrays
    names = FIELDS + ("x", "y")
    out["checksum"] = np.array([[float(np.sum(getattr(f, k))) for k in names] for f in paths])
    out["last"] = np.array([[getattr(f, k)[-1] for k in names] for f in paths])
    sample = list(range(0, len(paths), sample_stride))
    out["sample_idx"] = np.array(sample, dtype=np.int32)
    pk = pack_paths([paths[c] for c in sample], with_xy=True, nmax=max(len(f.t) for f in paths))
    for k in names:
        out[f"sample_{k}"] = pk[k]
    return out


def _obstacles_xy(spl, obs_s, obs_d):
    return np.array([spl.compute_pt(s) + ref_static.compute_ortogonal_vect(spl, s) * d for s, d in zip(obs_s, obs_d)])


def gen_dense_config4():
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    import otg_b200  # noqa: F401  (workload parameters only - plain NumPy draws)
    from otg_b200 import workloads
    sc = workloads.config4_scenario()
    spl = SplineTrajectory2D(SPLINE_WX, SPLINE_WY)
    obs_xy = _obstacles_xy(spl, sc["obs_s"], sc["obs_d"])
    t = time.time()
    out = _dense_reference((sc["p0"], sc["s0"], sc["t0"], sc["dd"], obs_xy, SPLINE_WX, SPLINE_WY, 97))
    print(f"  dense config 4: {len(out['nsteps'])} candidates, {int(out['nsteps'].sum())} cand-ts, argmin "
          f"{int(out['argmin'])}, masked {int(out['argmin_masked'])}, kept {int(out['keep'].sum())}, {time.time() - t:.1f} s")
    save("dense_config4", p0=np.array(sc["p0"]), s0=np.array(sc["s0"]), t0=np.array(sc["t0"]), dd=np.array(sc["dd"]),
         obs_s=sc["obs_s"], obs_d=sc["obs_d"], obs_xy=obs_xy, **out)


def gen_config5_sample(n_sample=16):
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    import otg_b200  # noqa: F401
    from otg_b200 import workloads
    sc = workloads.config5_scenarios()
    spl = SplineTrajectory2D(SPLINE_WX, SPLINE_WY)
    B = sc["p0"].shape[0]
    picks = [int(round(i * (B - 1) / (n_sample - 1))) for i in range(n_sample)]
    jobs = []
    obs_all = []
    for b in picks:
        fs, fd = workloads.moving_obstacle_frenet(sc["obs_s"][b], sc["obs_d"][b], sc["obs_speed"][b],
                                                  sc["obs_offset"][b], sc["t0"][b])
        obs_xy = _obstacles_xy(spl, fs, fd)
        obs_all.append(obs_xy)
        jobs.append((sc["p0"][b], sc["s0"][b], float(sc["t0"][b]), float(sc["dd"][b]), obs_xy, SPLINE_WX, SPLINE_WY, 997))
    t = time.time()
    with mp.Pool(min(len(jobs), os.cpu_count())) as pool:
        res = pool.map(_dense_reference, jobs)
    print(f"  config 5 sample: {n_sample} scenarios in {time.time() - t:.1f} s")
    out = dict(scenario_idx=np.array(picks, dtype=np.int32), obs_xy=np.array(obs_all))
    cmax = max(len(r["nsteps"]) for r in res)
    for key, fill, dtype in (("nsteps", 0, np.int32), ("cd", np.nan, np.float64), ("ctot", np.nan, np.float64),
                             ("keep", 0, np.uint8), ("first_blocker", -1, np.int16)):
        arr = np.full((n_sample, cmax), fill, dtype=dtype)
        for i, r in enumerate(res):
            arr[i, :len(r[key])] = r[key]
        out[key] = arr
    out["ncand"] = np.array([len(r["nsteps"]) for r in res], dtype=np.int32)
    out["argmin"] = np.array([int(r["argmin"]) for r in res], dtype=np.int32)
    out["argmin_masked"] = np.array([int(r["argmin_masked"]) for r in res], dtype=np.int32)
    out["dsi_interval"] = np.array([r["dsi_interval"] for r in res])
    out["cv_row"] = np.full((n_sample, cmax // 80), np.nan)
    for i, r in enumerate(res):
        out["cv_row"][i, :len(r["cv"]) // 80] = r["cv"][::80]
    for i, r in enumerate(res):
        print(f"    scenario {picks[i]}: {out['ncand'][i]} cands, kept {int(r['keep'].sum())}, argmin {int(r['argmin'])}"
              f" masked {int(r['argmin_masked'])}")
    save("config5_sample", **out)


# ------------------------------------------------------------------------------------------------
def gen_optimal_targets():
    """The open-loop flows of lib/tests/test_optimal_frenet.py:60-102 (stop / merge / follow targets)."""
    out = {}
    targets = {
        "stop": Target(s0=op.s0_t, s1=op.s1_t, Ts=op.Ts).get_frenet_stop_target(),
        "merge": Target(s0=op.s0_t, s1=op.s1_t, Ts=op.Ts).get_frenet_merge_target(Target(s0=op.s0_tb, s1=op.s1_tb, Ts=op.Ts)),
        "follow": Target(s0=op.s0_t, s1=op.s1_t, Ts=op.Ts).get_frenet_follow_target(D0=2),
    }
    for name, tg in targets.items():
        for k in ("t", "s", "dot_s", "ddot_s", "dddot_s"):
            out[f"{name}_target_{k}"] = np.array(getattr(tg, k), dtype=np.float64)
        pl = TrajectoryPlannerOptimal(TrajectoryPlannerParamsOptimal())
        pl.initialize(t0=op.t0, p0=op.p, s0=op.s, s_target=tg)
        pk = pack_paths(pl.paths)
        for k, v in pk.items():
            out[f"{name}_init_{k}"] = v
        out[f"{name}_init_argmin_ct"] = np.array(pl.paths.index(pl.opt_path_ct))
        out[f"{name}_init_argmin_ctot"] = np.array(pl.paths.index(pl.opt_path_tot))
        for j, t in enumerate(op.Tn):
            pl.replan_ct(t)
            pk = pack_paths(pl.paths)
            for k in ("nsteps", "cd", "ct", "ctot", "s", "d"):
                out[f"{name}_ct{j}_{k}"] = pk[k]
            out[f"{name}_ct{j}_p0"] = np.array(pl.p0, dtype=np.float64)
            out[f"{name}_ct{j}_s0"] = np.array(pl.s0, dtype=np.float64)
            out[f"{name}_ct{j}_argmin_ct"] = np.array(pl.paths.index(pl.opt_path_ct))
    # cd/cv replans without target + the attribute-mutation flow (:98-102)
    pl = TrajectoryPlannerOptimal(TrajectoryPlannerParamsOptimal())
    pl.initialize(t0=op.t0, p0=op.p, s0=op.s)
    for j, t in enumerate(op.Tn):
        pl.replan_cd_cv(t)
        pk = pack_paths(pl.paths)
        for k in ("nsteps", "cd", "cv", "ctot"):
            out[f"cdcv{j}_{k}"] = pk[k]
        out[f"cdcv{j}_p0"] = np.array(pl.p0, dtype=np.float64)
        out[f"cdcv{j}_s0"] = np.array(pl.s0, dtype=np.float64)
        out[f"cdcv{j}_argmin_cd"] = np.array(pl.paths.index(pl.opt_path_d))
        out[f"cdcv{j}_argmin_cv"] = np.array(pl.paths.index(pl.opt_path_s))
    pl.initialize(t0=op.t0, p0=op.px, s0=op.sx)
    pl.klat = 0.3
    pl.t_interval = (pl.min_t, pl.max_t, pl.dt / 2)
    pl.di_interval = (-1, 2.45, 0.2)
    ret = pl.replanner(time=0)
    pk = pack_paths(pl.paths)
    for k, v in pk.items():
        out[f"mutated_{k}"] = v
    out["mutated_ret"] = np.array([ret[0], ret[1]], dtype=np.float64)
    out["mutated_argmin"] = np.array(pl.paths.index(ret[2]))
    out["mutated_p0"] = np.array(pl.p0, dtype=np.float64)
    out["mutated_s0"] = np.array(pl.s0, dtype=np.float64)
    save("optimal_targets", **out)


def gen_follow_mode():
    """V1 follow mode with an ObstacleTrajectory target (trajectory_planner.py:131-144), the branch the
    moving-obstacle driver takes when every candidate is blocked (test_moving_obstacle_planner.py:92-97)."""
    trajectory = CircleTrajectory2D(8, 5, 2)
    transformer = FrenetGNTransformOld(0.5, 10)
    robot_p = np.array([3.0, 0.2, 0.05])
    transformer.estimatePosition(trajectory, robot_p)
    tgt = ObstacleTrajectory(o=-1, s=0.6, t0=7, t=trajectory)
    tgt.set_transformer(transformer)
    pl = TrajectoryPlannerV1(TrajectoryPlannerParams())
    pl.initialize(t0=0.1, p0=(0.2, 0.0, 0.0), s0=(3.0, 1.0, 0.0))
    out = {}
    for j, t in enumerate((0.2, 0.3, 0.4)):
        ret = pl.replanner(t, s_target=tgt)
        T = np.arange(*pl.t_interval)
        trip = []
        for tj in T:
            tt = pl.t_initial + tj
            trip.append((pl.s0[0] + tgt.compute_s(tt) - pl.d_target + pl.delta_t * tgt.compute_ds(tt),
                         pl.s0[1] + tgt.compute_ds(tt) - pl.delta_t * tgt.compute_dds(tt),
                         pl.s0[2] + tgt.compute_dds(tt)))
        pk = pack_paths(pl.paths)
        for k, v in pk.items():
            out[f"r{j}_{k}"] = v
        out[f"r{j}_triples"] = np.array(trip, dtype=np.float64)
        out[f"r{j}_time"] = np.array(t)
        out[f"r{j}_p0"] = np.array(pl.p0, dtype=np.float64)
        out[f"r{j}_s0"] = np.array(pl.s0, dtype=np.float64)
        out[f"r{j}_si_interval"] = np.array(pl.si_interval, dtype=np.float64)
        out[f"r{j}_argmin"] = np.array(pl.paths.index(pl.opt_path_tot))
        out[f"r{j}_target_s_ds_dds"] = np.array([[tgt.compute_s(pl.t_initial + tj), tgt.compute_ds(pl.t_initial + tj),
                                                  tgt.compute_dds(pl.t_initial + tj)] for tj in T])
    save("follow_mode", **out)


def gen_sensor_gate():
    rng = np.random.default_rng(11)
    robot = Unicycle()
    poses, obs, masks = [], [], []
    for _ in range(64):
        pose = np.array([rng.uniform(-2, 10), rng.uniform(-2, 10), rng.uniform(-np.pi, np.pi)])
        robot.set_initial_pose(pose.copy())
        pts = pose[:2] + rng.normal(0, 2.5, (16, 2))
        lst = [StaticObstacle(p) for p in pts]
        sensor = ProximitySensor(3.5, np.pi / 6)
        sensor.attach(robot)
        sensor.set_obstacle(lst)
        found = {id(o) for o in sensor.sense(pose)}
        poses.append(pose)
        obs.append(pts)
        masks.append([1 if id(o) in found else 0 for o in lst])
    save("sensor_gate", pose=np.array(poses), obs_xy=np.array(obs), mask=np.array(masks, dtype=np.uint8),
         range=np.array(3.5), aperture=np.array(np.pi / 6))


def gen_dt_sequences():
    """Replan gating of the DT variants (trajectory_planner_dt.py:229-233, trajectory_planner_obstacles.py:238-242):
    a run of `replanner(t)` calls; records the returned states and whether a new lattice was generated."""
    out = {}
    for name, cls, pcls in (("dt", TrajectoryPlannerV1DT, TrajectoryPlannerParamsDT),
                            ("dt_obstacles", TrajectoryPlannerV1DTObstacles, TrajectoryPlannerParamsDTObstacles)):
        pl = cls(pcls())
        pl.initialize(t0=0.1, p0=(0.05, 0.0, 0.0), s0=(1.0, 0.3, 0.0), dd=0.1)
        times = np.arange(0.1, 4.0, 0.1)
        ret_s, ret_p, replanned, argmin, nc = [], [], [], [], []
        for i, t in enumerate(times):
            before = pl.opt_path_tot
            kw = dict(force=True) if (name == "dt_obstacles" and i == 20) else {}
            s_, p_ = pl.replanner(t, **kw)
            ret_s.append(s_)
            ret_p.append(p_)
            replanned.append(pl.opt_path_tot is not before)
            argmin.append(pl.paths.index(pl.opt_path_tot))
            nc.append(len(pl.paths))
        out[f"{name}_times"] = times
        out[f"{name}_ret_s"] = np.array(ret_s, dtype=np.float64)
        out[f"{name}_ret_p"] = np.array(ret_p, dtype=np.float64)
        out[f"{name}_replanned"] = np.array(replanned, dtype=np.uint8)
        out[f"{name}_argmin"] = np.array(argmin, dtype=np.int32)
        out[f"{name}_ncand"] = np.array(nc, dtype=np.int32)
        print(f"  {name}: replanned {int(np.sum(replanned))} of {len(times)} calls")
    save("dt_sequences", **out)


def gen_follow_fallback():
    """The branch of the moving-obstacle driver that the canned run never takes (test_moving_obstacle_planner.py:92-97):
    every candidate is blocked, so the planner is re-run in follow mode behind the first blocking obstacle.  A slow
    obstacle is parked on the track just ahead of the robot; the reference driver loop is run for a few steps."""
    t_vect, robot, trajectory, transformer, controller, planner = _sim_objects()
    np.random.seed(11)
    ob = MovingObstacle(np.array([1.5, 0.0]), trajectory)
    # the lattice is started at arc length 0.5 (below); the obstacle sits 0.5 m further on - inside the robot radius of the
    # point every candidate starts from - and moves on at 0.3 m/s
    ob.trajectory.offset, ob.trajectory.speed, ob.trajectory.time_offset = 0, 0.3, 1.0
    ob.step(0)
    sensor = ProximitySensor(3.5, np.pi / 6)
    sensor.attach(robot)
    sensor.set_obstacle([ob])
    n = 12
    rec = dict(time=t_vect[:n].copy(), fired=np.zeros(n, dtype=np.uint8), ret_s=np.zeros((n, 3)), ret_p=np.zeros((n, 3)),
               ncand=np.zeros(n, dtype=np.int32), argmin=np.zeros(n, dtype=np.int32), ctot_min=np.zeros(n),
               robot_pose=np.zeros((n, 3)), si=np.zeros((n, 3)), obs_xy=np.zeros((n, 2)))
    robot_p = robot.p
    transformer.estimatePosition(trajectory, robot_p)
    f = transformer.transform(robot_p)
    planner.initialize(t0=t_vect[0], p0=(f[1], 0.0, 0.0), s0=(0.5, 0.0, 0.0))
    for i in range(n):
        est_pt = transformer.estimatePosition(trajectory, robot_p)
        robot_fpose = transformer.transform(robot_p)
        pos_s, pos_d = planner.replanner(t_vect[i])
        paths = ref_moving.frenet_to_glob(planner, trajectory, planner.paths)
        sensor.step_obstacle(t_vect[i])
        rec["obs_xy"][i] = ob.position
        planner.paths, leading = ref_moving.check_paths(paths, sensor, robot_p)
        if len(planner.paths) == 0:
            rec["fired"][i] = 1
            s_target = leading[0].trajectory
            s_target.set_transformer(transformer)
            pos_s, pos_d = planner.replanner(t_vect[i], s_target=s_target)
            ref_moving.frenet_to_glob(planner, trajectory, planner.paths)
        planner.opt_path_tot = min(planner.paths, key=attrgetter('ctot'))
        rec["ret_s"][i], rec["ret_p"][i] = pos_s, pos_d
        rec["ncand"][i] = len(planner.paths)
        rec["argmin"][i] = planner.paths.index(planner.opt_path_tot)
        rec["ctot_min"][i] = planner.opt_path_tot.ctot
        rec["si"][i] = planner.si_interval
        error = np.array([0, pos_d[0]]) - robot_fpose[0:2]
        derror = np.array([pos_s[1], pos_d[1]])
        u = controller.compute(robot_fpose, error, derror, trajectory.compute_curvature(est_pt))
        robot_p, _ = robot.step(u, 0.1)
        rec["robot_pose"][i] = robot_p
    print(f"  follow fallback fired in {int(rec['fired'].sum())} of {n} steps; candidates {rec['ncand'].tolist()}")
    save("follow_fallback", **rec)


def gen_dt_lane():
    """Duckietown lane run pieces with the reference's own functions
    (lib/tests/test_mapper_planner_obstacles.py:25-135): V1DTObstacles lattice, parabola path built like
    MapperSemanticObstacles.build_trajectory (lib/mapper/mapper_obstacles.py:169-186), projection-relative
    frenet_to_glob_planner, check_paths with off-road parabolas and two obstacle point sets.  The camera -> robot
    mapping is perception (out of scope): obstacles are given in the robot frame through an identity `cam2rob`."""
    import lib.tests.test_mapper_planner_obstacles as dtm
    from lib.trajectory import Trajectory

    class IdentityMapper:
        def cam2rob(self, pts):
            return np.asarray(pts, dtype=np.float64)

    out = {}
    cases = [  # (a, b, c) of the lane centre, right / left boundary fits (or None), projection, state, obstacles (r, lat)
        dict(fit=(0.3, 0.05, 0.01), rw=(0.32, 0.05, -0.29), lw=(0.28, 0.05, 0.33), proj=0.07, s0=(0.0, 0.2, 0.0), p0=(0.01, 0.0, 0.0),
             obs=[((0.60, 0.05), (0.58, 0.12))]),
        dict(fit=(-0.4, 0.3, -0.02), rw=None, lw=(-0.4, 0.3, 0.25), proj=0.05, s0=(0.03, 0.45, 0.1), p0=(-0.03, 0.02, 0.0),
             obs=[((0.50, -0.27), (0.48, -0.21)), ((0.95, -0.45), (0.92, -0.38))]),
        dict(fit=(0.3, -0.1, 0.0), rw=(0.3, -0.1, -0.22), lw=None, proj=0.1, s0=(0.0, 0.0, 0.0), p0=(0.0, 0.0, 0.0), obs=[]),
    ]
    out["n_cases"] = np.array(len(cases))
    for i, cs in enumerate(cases):
        a, b, c = cs["fit"]
        trajectory = Trajectory()
        trajectory.compute_pt = lambda x, a=a, b=b, c=c: np.array([x, a * x ** 2 + b * x + c])
        trajectory.compute_first_derivative = lambda x, a=a, b=b: np.array([1, 2 * a * x + b])
        pl = TrajectoryPlannerV1DTObstacles(TrajectoryPlannerParamsDTObstacles())
        pl.initialize(t0=0, p0=cs["p0"], s0=cs["s0"])
        pl.replanner(time=1 / 30, force=True)
        paths = dtm.frenet_to_glob_planner(pl, trajectory, pl.paths, cs["proj"])
        obstacles = [{"end_point": r, "center": r, "end_right": r, "end_lat": l, "end_top": l} for r, l in cs["obs"]]
        rw = None if cs["rw"] is None else np.array(cs["rw"])
        lw = None if cs["lw"] is None else np.array(cs["lw"])
        if not obstacles:
            kept = dtm.check_paths(paths, obstacles, np.array([0.07, 0.0, 0.0]), IdentityMapper(), rw, lw)
        else:
            # check_paths (:112-135) tests `measure_obst != []` on an ndarray, which NumPy >= 2 refuses to broadcast (older
            # NumPy warned and took the branch).  Same composition from the reference's own piece functions, branch taken:
            _, _, obs_r, obs_lat, _ = dtm.obstacles_coordinates(obstacles, IdentityMapper())
            kept = []
            for f in paths:
                if rw is not None and dtm.check_offroad(f, dtm.get_vertex_and_focus_distance(rw), 'right'):
                    continue
                if lw is not None and dtm.check_offroad(f, dtm.get_vertex_and_focus_distance(lw), 'left'):
                    continue
                if not dtm.check_collisions(f, obs_r) or not dtm.check_collisions(f, obs_lat):
                    continue
                kept.append(f)
        kept_ids = {id(f) for f in kept}
        pk = pack_paths(paths, with_xy=True)
        for k, v in pk.items():
            out[f"c{i}_{k}"] = v
        out[f"c{i}_keep"] = np.array([1 if id(f) in kept_ids else 0 for f in paths], dtype=np.uint8)
        out[f"c{i}_offroad_right"] = np.array([0 if rw is None else int(dtm.check_offroad(f, dtm.get_vertex_and_focus_distance(rw), 'right'))
                                               for f in paths], dtype=np.uint8)
        out[f"c{i}_offroad_left"] = np.array([0 if lw is None else int(dtm.check_offroad(f, dtm.get_vertex_and_focus_distance(lw), 'left'))
                                              for f in paths], dtype=np.uint8)
        out[f"c{i}_fit"] = np.array(cs["fit"], dtype=np.float64)
        out[f"c{i}_rw"] = np.array(cs["rw"] if cs["rw"] else [np.nan] * 3, dtype=np.float64)
        out[f"c{i}_lw"] = np.array(cs["lw"] if cs["lw"] else [np.nan] * 3, dtype=np.float64)
        out[f"c{i}_proj"] = np.array(cs["proj"])
        out[f"c{i}_init_p0"] = np.array(cs["p0"], dtype=np.float64)
        out[f"c{i}_init_s0"] = np.array(cs["s0"], dtype=np.float64)
        out[f"c{i}_p0"] = np.array(pl.p0, dtype=np.float64)          # state after the forced replan = the lattice's start state
        out[f"c{i}_s0"] = np.array(pl.s0, dtype=np.float64)
        out[f"c{i}_obs_r"] = np.array([r for r, _ in cs["obs"]], dtype=np.float64).reshape(-1, 2)
        out[f"c{i}_obs_lat"] = np.array([l for _, l in cs["obs"]], dtype=np.float64).reshape(-1, 2)
        out[f"c{i}_argmin"] = np.array(pl.paths.index(pl.opt_path_tot))
        out[f"c{i}_argmin_kept"] = np.array(paths.index(min(kept, key=attrgetter('ctot'))) if kept else -1)
        win = dtm.frenet_to_glob(trajectory, pl, cs["proj"])
        out[f"c{i}_winner_xy"] = win
        out[f"c{i}_normal"] = np.array([dtm.compute_ortogonal_vect(trajectory, s) for s in (0.0, 0.1, 0.37, -0.05)])
        print(f"  dt lane case {i}: {len(paths)} candidates, kept {len(kept)}, off-road right {int(out[f'c{i}_offroad_right'].sum())} "
              f"left {int(out[f'c{i}_offroad_left'].sum())}")
    out["agent_safety_rad"] = np.array(dtm.dp.AGENT_SAFETY_RAD)
    save("dt_lane", **out)


GENERATORS = {
    "kat_variants": gen_kat_variants,
    "paths": gen_paths,
    "sim_full": lambda: gen_sim("full"),
    "sim_static": lambda: gen_sim("static"),
    "sim_moving": lambda: gen_sim("moving"),
    "optimal_targets": gen_optimal_targets,
    "follow_mode": gen_follow_mode,
    "sensor_gate": gen_sensor_gate,
    "dense_config4": gen_dense_config4,
    "config5_sample": gen_config5_sample,
    "dt_sequences": gen_dt_sequences,
    "follow_fallback": gen_follow_fallback,
    "dt_lane": gen_dt_lane,
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    for name, fn in GENERATORS.items():
        if args.only and name not in args.only:
            continue
        print(f"[{name}]")
        t0 = time.time()
        fn()
        print(f"  {time.time() - t0:.1f} s")
