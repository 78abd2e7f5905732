"""GPU tests of the drop-in API: planner classes, driver functions and the three closed-loop experiments,
against fixtures recorded from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from _golden import golden, prefixed, assert_close, near_tie
from oracle import frenet_oracle as fo

pytestmark = pytest.mark.gpu
RT = 1e-9
NAMES9 = ("t", "d", "dot_d", "ddot_d", "dddot_d", "s", "dot_s", "ddot_s", "dddot_s")


def _classes():
    import otg_b200  # noqa: F401
    from otg_b200 import planner as P
    return {"v1": (P.TrajectoryPlannerV1, P.TrajectoryPlannerParams), "dt": (P.TrajectoryPlannerV1DT, P.TrajectoryPlannerParamsDT),
            "dt_obstacles": (P.TrajectoryPlannerV1DTObstacles, P.TrajectoryPlannerParamsDTObstacles),
            "optimal": (P.TrajectoryPlannerOptimal, P.TrajectoryPlannerParamsOptimal)}


def _check_frenet(f, g, c, what):
    n = int(g["nsteps"][c])
    assert len(f.t) == n
    for name in NAMES9:
        assert_close(np.array(getattr(f, name)), g[name][c, :n], RT, f"{what}.{name}")
    for k in ("cd", "cv", "ct", "ctot"):
        assert getattr(f, k) == pytest.approx(float(g[k][c]), rel=RT, abs=1e-300)


def test_initialize_matches_reference_for_all_variants():
    z = golden("kat_variants")
    cls = _classes()
    for i in range(int(z["n_cases"])):
        g = prefixed(z, f"c{i}_")
        K, Pm = cls[str(g["variant"])]
        pl = K(Pm())
        pl.initialize(t0=float(g["t0"]), p0=tuple(g["p0"]), s0=tuple(g["s0"]), dd=float(g["dd"]))
        assert tuple(pl.di_interval) == tuple(g["di_interval"]) and tuple(pl.dsi_interval) == tuple(g["dsi_interval"])
        assert np.allclose(pl.t_interval, g["t_interval"], rtol=0, atol=0)
        assert len(pl.paths) == len(g["nsteps"])
        for key, attr in (("ctot", "opt_path_tot"), ("cd", "opt_path_d"), ("cv", "opt_path_s")):
            w = getattr(pl, attr)
            ref = int(g[f"argmin_{key}"])
            got = pl.paths.index(w)
            assert got == ref or near_tie(g[key], got, ref), (i, key)
            if got == ref:
                _check_frenet(w, g, ref, f"kat{i}.{attr}")
        assert pl.opt_path_cd is pl.opt_path_d and pl.opt_path_cv is pl.opt_path_s
        # element access on the lazy candidate list
        C = len(pl.paths)
        for c in (0, C // 3, C - 1):
            _check_frenet(pl.paths[c], g, c, f"kat{i}.paths[{c}]")
        assert [p.index for p in pl.paths[:3]] == [pl.paths[0].index, pl.paths[1].index, pl.paths[2].index]
        from operator import attrgetter
        assert min(pl.paths, key=attrgetter("ctot")).ctot == pl.opt_path_tot.ctot     # the reference idiom still works


def test_assert_and_attribute_semantics():
    K, Pm = _classes()["v1"]
    pl = K(Pm(), kj=0.002)                 # kwargs override the params object
    assert pl.kj == 0.002 and pl.klat == 1
    with pytest.raises(AssertionError):
        pl.initialize(t0=0.0, p0=(0, 0, 0), s0=(0, 0, 0), dd=10.0)
    pl = K(**Pm().dict())                  # pure keyword construction
    pl.initialize(t0=0.1, p0=(0, 0, 0), s0=(0, 0, 0))
    assert len(pl.paths) == 320 and pl.paths.index(pl.opt_path_tot) == 312
    assert pl.opt_path_tot.ctot == pytest.approx(0.33450470400000004, rel=RT)
    s0, p0 = pl.replanner(0.2)
    assert s0 == pl.s0 and p0 == pl.p0
    old = pl.paths                         # never looked at: still only in HBM
    pl.replanner(0.25)
    with pytest.raises(RuntimeError):      # the previous lattice's device buffers were reused
        old[0]
    kept = pl.paths.materialize()
    pl.replanner(0.3)
    assert len(kept[5].t) > 0              # materialised views survive a replan
    pl.paths = [kept[0]]                   # drivers assign plain lists
    pl.opt_path_tot = kept[0]
    k = round((0.25 - kept[0].t[0]) / pl.delta_t)
    assert pl.optimal_at_time(0.25, pl.opt_path_tot, "s") == (kept[0].s[k], kept[0].dot_s[k], kept[0].ddot_s[k])
    assert pl.optimal_at_time(-1.0, pl.opt_path_tot, "d")[0] == kept[0].d[0]        # clamped below
    assert pl.optimal_at_time(99.0, pl.opt_path_tot, "d")[0] == kept[0].d[-1]       # clamped above


def test_dt_replan_gating():
    z = golden("dt_sequences")
    cls = _classes()
    for name in ("dt", "dt_obstacles"):
        K, Pm = cls[name]
        pl = K(Pm())
        pl.initialize(t0=0.1, p0=(0.05, 0.0, 0.0), s0=(1.0, 0.3, 0.0), dd=0.1)
        for i, t in enumerate(z[f"{name}_times"]):
            before = pl.opt_path_tot
            kw = dict(force=True) if (name == "dt_obstacles" and i == 20) else {}
            s_, p_ = pl.replanner(t, **kw)
            assert_close(np.array(s_), z[f"{name}_ret_s"][i], RT, f"{name}[{i}].s", atol_scale=1.0)
            assert_close(np.array(p_), z[f"{name}_ret_p"][i], RT, f"{name}[{i}].p", atol_scale=1.0)
            assert (pl.opt_path_tot is not before) == bool(z[f"{name}_replanned"][i]), (name, i)
            assert len(pl.paths) == int(z[f"{name}_ncand"][i])
            assert pl.paths.index(pl.opt_path_tot) == int(z[f"{name}_argmin"][i])


def test_optimal_planner_flows():
    """lib/tests/test_optimal_frenet.py:60-102 headless: stop / merge / follow targets, cd-cv replans, attribute mutation."""
    import otg_b200  # noqa: F401
    from otg_b200.optimal_frenet_frame import Target, OptimalParameters as op
    z = golden("optimal_targets")
    K, Pm = _classes()["optimal"]
    targets = {
        "stop": Target(s0=op.s0_t, s1=op.s1_t, Ts=op.Ts).get_frenet_stop_target(),
        "merge": Target(s0=op.s0_t, s1=op.s1_t, Ts=op.Ts).get_frenet_merge_target(Target(s0=op.s0_tb, s1=op.s1_tb, Ts=op.Ts)),
        "follow": Target(s0=op.s0_t, s1=op.s1_t, Ts=op.Ts).get_frenet_follow_target(D0=2),
    }
    for name, tg in targets.items():
        assert np.array_equal(np.array(tg.s), z[f"{name}_target_s"])
        pl = K(Pm())
        pl.initialize(t0=op.t0, p0=op.p, s0=op.s, s_target=tg)
        g = prefixed(z, f"{name}_init_")
        assert len(pl.paths) == len(g["nsteps"])
        assert pl.paths.index(pl.opt_path_ct) == int(g["argmin_ct"]) and pl.paths.index(pl.opt_path_tot) == int(g["argmin_ctot"])
        for j, t in enumerate(op.Tn):
            pl.replan_ct(t)
            g = prefixed(z, f"{name}_ct{j}_")
            assert_close(np.array(pl.p0), g["p0"], RT, f"{name}.ct{j}.p0", atol_scale=1.0)
            assert_close(np.array(pl.s0), g["s0"], RT, f"{name}.ct{j}.s0", atol_scale=1.0)
            assert len(pl.paths) == len(g["nsteps"]) and pl.paths.index(pl.opt_path_ct) == int(g["argmin_ct"])
            assert_close(np.array([p.ct for p in pl.paths]), g["ct"], RT, f"{name}.ct{j}.ct")
    pl = K(Pm())
    pl.initialize(t0=op.t0, p0=op.p, s0=op.s)
    for j, t in enumerate(op.Tn):
        pl.replan_cd_cv(t)
        g = prefixed(z, f"cdcv{j}_")
        assert_close(np.array(pl.p0), g["p0"], RT, f"cdcv{j}.p0", atol_scale=1.0)
        assert_close(np.array(pl.s0), g["s0"], RT, f"cdcv{j}.s0", atol_scale=1.0)
        thr = pl.low_speed_threshold
        if abs(g["s0"][1] - thr) <= 1e-9 * thr and (pl.s0[1] < thr) != (g["s0"][1] < thr):
            break      # speed exactly on the low-speed threshold and the last bit differs: the north star's tie exception
        assert len(pl.paths) == len(g["nsteps"])
        assert pl.paths.index(pl.opt_path_d) == int(g["argmin_cd"]) and pl.paths.index(pl.opt_path_s) == int(g["argmin_cv"])
    pl.initialize(t0=op.t0, p0=op.px, s0=op.sx)
    pl.klat = 0.3
    pl.t_interval = (pl.min_t, pl.max_t, pl.dt / 2)
    pl.di_interval = (-1, 2.45, 0.2)
    d0, s0, best = pl.replanner(time=0)
    g = prefixed(z, "mutated_")
    assert_close(np.array([d0, s0]), g["ret"], RT, "mutated.ret", atol_scale=1.0)
    assert len(pl.paths) == len(g["nsteps"]) and pl.paths.index(best) == int(g["argmin"])
    _check_frenet(best, g, int(g["argmin"]), "mutated.best")


def test_follow_mode_with_obstacle_trajectory_target():
    """V1 follow mode (trajectory_planner.py:131-144) behind an ObstacleTrajectory, as the moving-obstacle driver
    does when every candidate is blocked."""
    import otg_b200  # noqa: F401
    from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
    from otg_b200.sensor import ObstacleTrajectory
    from otg_b200.sim import FrenetGNTransformOld
    from otg_b200.trajectory import CircleTrajectory2D
    z = golden("follow_mode")
    trajectory = CircleTrajectory2D(8, 5, 2)
    transformer = FrenetGNTransformOld(0.5, 10)
    transformer.estimatePosition(trajectory, np.array([3.0, 0.2, 0.05]))
    tgt = ObstacleTrajectory(o=-1, s=0.6, t0=7, t=trajectory)
    tgt.set_transformer(transformer)
    pl = TrajectoryPlannerV1(TrajectoryPlannerParams())
    pl.initialize(t0=0.1, p0=(0.2, 0.0, 0.0), s0=(3.0, 1.0, 0.0))
    for j, t in enumerate((0.2, 0.3, 0.4)):
        pl.replanner(t, s_target=tgt)
        g = prefixed(z, f"r{j}_")
        assert_close(np.array(pl.p0), g["p0"], RT, f"follow{j}.p0", atol_scale=1.0)
        assert_close(np.array(pl.s0), g["s0"], RT, f"follow{j}.s0", atol_scale=1.0)
        assert tuple(pl.si_interval) == tuple(g["si_interval"])
        assert len(pl.paths) == len(g["nsteps"]) and pl.paths.index(pl.opt_path_tot) == int(g["argmin"])
        _check_frenet(pl.paths[17], g, 17, f"follow{j}.paths[17]")


@pytest.mark.parametrize("kind", ["full", "static", "moving"])
def test_closed_loop_experiments_match_reference(kind):
    """tester.py -t planner_full | planner_obstacles | planner_moving_obstacles on the circle configuration, headless:
    every step's chosen candidate and the whole robot trajectory against the recorded reference run."""
    import otg_b200  # noqa: F401
    from otg_b200 import drivers
    z = golden(f"sim_planner_{kind}")
    picked, ncand = [], []

    def hook(i, planner):
        picked.append(planner.opt_path_tot.index)
        ncand.append(len(planner.paths))

    fn = {"full": drivers.planner_full, "static": drivers.obstacle_planner, "moving": drivers.moving_obstacle_planner}[kind]
    el = fn.default_elements()
    if kind == "full":
        res = fn.simulate(el.t_vect, el.trajectory, el.robot, el.transformer, el.controller, el.planner, hook=hook)
    else:
        from otg_b200.sensor import ProximitySensor, StaticObstacle, MovingObstacle
        pts = drivers.obstacle_planner.seeded_obstacles(el.t_vect, el.trajectory)
        assert_close(pts.T, z["obs_start"], 1e-12, "seeded obstacles", atol_scale=20.0)
        obs = [StaticObstacle(pts[:, i]) if kind == "static" else MovingObstacle(pts[:, i], el.trajectory) for i in range(10)]
        sensor = ProximitySensor(3.5, np.pi / 6)
        sensor.attach(el.robot)
        sensor.set_obstacle(obs)
        res = fn.simulate(el.t_vect, el.trajectory, el.robot, el.transformer, el.controller, el.planner, sensor, hook=hook)
        assert res.follow_steps == []
    ref_pick = z["argmin"] if kind == "full" else z["argmin_masked"]
    picked = np.array(picked)
    bad = np.nonzero(picked != ref_pick)[0]
    m = len(ref_pick) if bad.size == 0 else int(bad[0])
    if m < len(ref_pick):
        # The only admissible divergence (north star: "exact except where a limit falls within 1e-6 relative of a
        # threshold"): the reference's longitudinal speed sits EXACTLY on low_speed_threshold (the lattice's integer
        # target speed 2 equals the threshold), so `ds0 < low_speed_threshold` is decided by the last bit of a
        # value whose rounding history differs between the reference's power-form evaluation and ours.  From
        # that replan on the two closed loops are different (equally valid) runs.
        thr = el.planner.low_speed_threshold
        assert kind == "moving" and m >= 290 and abs(z["s0"][m, 1] - thr) <= 1e-9 * thr, f"diverged at step {m}"
    assert np.array_equal(picked[:m], ref_pick[:m])
    if kind != "full":
        assert np.array_equal(np.array(ncand)[:m], z["keep"].sum(axis=1)[:m])
    assert_close(res.robot_pose[:m], z["robot_pose"][:m], 1e-7, f"{kind} robot trajectory", atol_scale=20.0)
    if m == len(ref_pick):
        known = {"full": [7.73155170389, 8.89763727946, 15.6348962265], "static": [7.95339244018, 8.00629868489, 16.2723448421],
                 "moving": [7.36882758358, 8.92013375494, 15.4653965574]}[kind]
        assert np.allclose(res.robot_pose[-1], known, rtol=0, atol=1e-6)


def test_driver_functions_on_plain_lists_and_all_blocked():
    """frenet_to_glob / check_paths / check_collisions handed ordinary lists of Frenet objects (not the planner's live
    lattice) go through the generic point / polyline kernels and agree with the lattice path; an all-blocked scene
    raises ValueError from the masked pick like the reference's min([])."""
    import copy
    import otg_b200  # noqa: F401
    from otg_b200 import drivers
    from otg_b200.drivers import obstacle_planner as st, moving_obstacle_planner as mv
    from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
    from otg_b200.sensor import ProximitySensor, StaticObstacle
    from otg_b200.sim import Unicycle
    from otg_b200.trajectory import CircleTrajectory2D
    tr = CircleTrajectory2D(8, 5, 2)
    robot = Unicycle()
    robot.set_initial_pose(np.array([3.0, 0.0, 0.0]))
    obstacles = [StaticObstacle(np.array(p)) for p in ([5.0, 0.3], [6.0, -1.0], [4.5, 2.5], [20.0, 20.0])]
    sensor = ProximitySensor(3.5, np.pi / 6)
    sensor.attach(robot)
    sensor.set_obstacle(obstacles)
    pl = TrajectoryPlannerV1(TrajectoryPlannerParams())
    pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(3.0, 2.0, 0.0))
    live = drivers.frenet_to_glob(pl, tr, pl.paths)
    plain = [copy.deepcopy(f) for f in live]
    for f in plain:
        f.x, f.y = [], []
    drivers.frenet_to_glob(pl, tr, plain)
    for a, b in zip(plain[::37], live[::37]):
        assert a.x == b.x and a.y == b.y
    kept_live, lead_live = mv.check_paths(live, sensor, robot.p)
    kept_plain, lead_plain = mv.check_paths(plain, sensor, robot.p)
    assert 0 < len(kept_live) < len(live)
    assert [f.index for f in kept_live] == [f.index for f in kept_plain]
    assert [id(o) for o in lead_live] == [id(o) for o in lead_plain] and len(lead_live) == len(live) - len(kept_live)
    assert len(st.check_paths(plain, sensor, robot.p)) == len(kept_plain)
    assert drivers.best_path(kept_live).index == drivers.best_path(kept_plain).index
    blocked = next(f for f in plain if f.index not in {k.index for k in kept_plain})
    ok, ob = mv.check_collisions(blocked, sensor.sense())
    assert ok is False and ob is lead_plain[0] or ob in lead_plain
    assert st.check_collisions(kept_plain[0], sensor.sense()) is True
    # oracle agrees
    L = fo.generate(fo.OracleParams.defaults("v1"), (0, 0, 0), (3.0, 2.0, 0.0), 0.1)
    x, y = fo.lattice_to_cartesian(fo.CirclePath(8, 5, 2), L)
    sensed = np.array([o.position for o in sensor.sense()])
    okm, _ = fo.check_collisions(x, y, L.nsteps, sensed)
    assert [f.index for f in kept_live] == np.nonzero(okm)[0].tolist()
    # nothing sensed -> everything passes untouched
    sensor.set_obstacle([obstacles[3]])
    assert st.check_paths(live, sensor, robot.p) is live
    # all blocked -> ValueError (static), empty + leading obstacles (moving)
    sensor.set_obstacle([StaticObstacle(np.array([3.0, 0.0]))])
    pl.replanner(0.2)
    paths = drivers.frenet_to_glob(pl, tr, pl.paths)
    kept, lead = mv.check_paths(paths, sensor, robot.p)
    assert len(kept) == 0 and len(lead) == len(paths)
    with pytest.raises(ValueError):
        drivers.best_path(kept)


def test_follow_mode_fallback_when_every_candidate_is_blocked():
    """The branch the canned moving-obstacle run never takes (test_moving_obstacle_planner.py:92-97): all candidates
    collide -> re-plan in follow mode behind the first blocking obstacle; recorded from the reference driver loop."""
    import otg_b200  # noqa: F401
    from otg_b200 import drivers
    from otg_b200.drivers import moving_obstacle_planner as mv
    from otg_b200.sensor import ProximitySensor, MovingObstacle
    z = golden("follow_fallback")
    el = mv.default_elements()
    np.random.seed(11)
    ob = MovingObstacle(np.array([1.5, 0.0]), el.trajectory)
    ob.trajectory.offset, ob.trajectory.speed, ob.trajectory.time_offset = 0, 0.3, 1.0
    ob.step(0)
    sensor = ProximitySensor(3.5, np.pi / 6)
    sensor.attach(el.robot)
    sensor.set_obstacle([ob])
    planner, trajectory, transformer, controller, robot = el.planner, el.trajectory, el.transformer, el.controller, el.robot
    robot_p = robot.p
    transformer.estimatePosition(trajectory, robot_p)
    f = transformer.transform(robot_p)
    planner.initialize(t0=el.t_vect[0], p0=(f[1], 0.0, 0.0), s0=(0.5, 0.0, 0.0))
    fired = []
    for i in range(len(z["time"])):
        t = el.t_vect[i]
        est_pt = transformer.estimatePosition(trajectory, robot_p)
        robot_fpose = transformer.transform(robot_p)
        pos_s, pos_d = planner.replanner(t)
        paths = drivers.frenet_to_glob(planner, trajectory, planner.paths)
        sensor.step_obstacle(t)
        assert_close(ob.position, z["obs_xy"][i], 1e-12, f"obstacle @ {i}", atol_scale=10.0)
        planner.paths, leading = mv.check_paths(paths, sensor, robot_p)
        if len(planner.paths) == 0:
            fired.append(i)
            assert len(leading) == 320 and all(o is ob for o in leading)
            s_target = leading[0].trajectory
            s_target.set_transformer(transformer)
            pos_s, pos_d = planner.replanner(t, s_target=s_target)
            drivers.frenet_to_glob(planner, trajectory, planner.paths)
        planner.opt_path_tot = drivers.best_path(planner.paths)
        assert len(planner.paths) == int(z["ncand"][i]), i
        assert planner.paths.index(planner.opt_path_tot) == int(z["argmin"][i]), i
        assert planner.opt_path_tot.ctot == pytest.approx(float(z["ctot_min"][i]), rel=RT)
        assert_close(np.array(pos_s), z["ret_s"][i], RT, f"ret_s @ {i}", atol_scale=1.0)
        assert_close(np.array(pos_d), z["ret_p"][i], RT, f"ret_p @ {i}", atol_scale=1.0)
        assert tuple(planner.si_interval) == tuple(z["si"][i])
        error = np.array([0, pos_d[0]]) - robot_fpose[0:2]
        derror = np.array([pos_s[1], pos_d[1]])
        u = controller.compute(robot_fpose, error, derror, trajectory.compute_curvature(est_pt))
        robot_p, _ = robot.step(u, 0.1)
        assert_close(robot_p, z["robot_pose"][i], 1e-9, f"pose @ {i}", atol_scale=10.0)
    assert fired == np.nonzero(z["fired"])[0].tolist() and len(fired) == 5


def test_more_obstacles_than_the_staging_capacity():
    """check_paths with more sensed obstacles than the planner's obstacle staging capacity falls back to the generic
    polyline kernel and gives the same survivors as the lattice kernel."""
    import otg_b200  # noqa: F401
    from otg_b200 import drivers
    from otg_b200.drivers import moving_obstacle_planner as mv
    from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
    from otg_b200.sensor import StaticObstacle
    from otg_b200.trajectory import CircleTrajectory2D
    tr = CircleTrajectory2D(8, 5, 2)
    rng = np.random.default_rng(5)
    obstacles = [StaticObstacle(np.array([rng.uniform(4, 12), rng.uniform(-3, 3)])) for _ in range(9)]

    class Sees:
        def sense(self, *a, **k):
            return obstacles

    out = []
    for cap in (4, 64):
        pl = TrajectoryPlannerV1(TrajectoryPlannerParams(), _n_obs_cap=cap)
        pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(3.0, 2.0, 0.0))
        paths = drivers.frenet_to_glob(pl, tr, pl.paths)
        kept, lead = mv.check_paths(paths, Sees(), None)
        out.append(([f.index for f in kept], [id(o) for o in lead], drivers.best_path(kept).index))
    assert out[0] == out[1] and 0 < len(out[0][0]) < 320


def test_empty_lattice_raises_like_min_of_empty_list():
    """d_d_s = 0.2, num_sample = 1 at ds0 = 0: both ends of the target-speed interval round to 0, np.arange is empty, the
    reference's initialize() dies in min([]) with ValueError - so does ours (no candidates, K5 reports EMPTY)."""
    import otg_b200  # noqa: F401
    from otg_b200 import _native as nat
    from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
    pl = TrajectoryPlannerV1(TrajectoryPlannerParams(d_d_s=0.2, num_sample=1))
    with pytest.raises(ValueError):
        pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(0.0, 0.0, 0.0))
    assert len(pl.paths) == 0 and pl._engine.result[0]["status"] == nat.STATUS_EMPTY
    pl2 = TrajectoryPlannerV1(TrajectoryPlannerParams(d_d_s=0.2, num_sample=1))
    pl2.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(0.0, 1.5, 0.0))       # a non-degenerate speed works with the same parameters
    assert len(pl2.paths) == len(np.arange(round(1.5 - 0.2), round(1.5 + 0.2 + 0.2), 0.2)) * 4 * 16 == 5 * 64


def test_duckietown_lane_filter_matches_reference():
    """projection-relative transform on a parabola lane path, off-road + two-point-set collision
    filter and cheapest survivor, against the reference's lib/tests/test_mapper_planner_obstacles.py functions."""
    from operator import attrgetter
    from otg_b200.drivers import dt_lane as dl
    from otg_b200.trajectory import ParabolaTrajectory2D

    class IdentityMapper:
        def cam2rob(self, pts):
            return np.asarray(pts, dtype=np.float64)

    z = golden("dt_lane")
    assert dl.AGENT_SAFETY_RAD == float(z["agent_safety_rad"])
    K, Pm = _classes()["dt_obstacles"]
    for i in range(int(z["n_cases"])):
        g = prefixed(z, f"c{i}_")
        lane = ParabolaTrajectory2D(*g["fit"])
        for s, ref in zip((0.0, 0.1, 0.37, -0.05), g["normal"]):
            assert_close(dl.compute_ortogonal_vect(lane, s), ref, 1e-15, "normal")
        pl = K(Pm())
        pl.initialize(t0=0, p0=tuple(g["init_p0"]), s0=tuple(g["init_s0"]))
        pl.replanner(time=1 / 30, force=True)
        assert_close(np.array(pl.s0), g["s0"], 1e-12, "s0")
        proj = float(g["proj"])
        paths = dl.frenet_to_glob_planner(pl, lane, pl.paths, proj)
        C = len(paths)
        assert C == len(g["nsteps"])
        for c in (0, C // 2, C - 1):
            n = int(g["nsteps"][c])
            assert_close(np.array(paths[c].x), g["x"][c, :n], 1e-12, "x", atol_scale=1.0)
            assert_close(np.array(paths[c].y), g["y"][c, :n], 1e-12, "y", atol_scale=1.0)
        obstacles = [{"end_point": r, "center": r, "end_right": r, "end_lat": l, "end_top": l}
                     for r, l in zip(g["obs_r"], g["obs_lat"])]
        rw = None if np.isnan(g["rw"]).any() else g["rw"]
        lw = None if np.isnan(g["lw"]).any() else g["lw"]
        kept = dl.check_paths(paths, obstacles, np.array([0.07, 0.0, 0.0]), IdentityMapper(), rw, lw)
        ref_keep = g["keep"].astype(bool)
        assert len(kept) == int(ref_keep.sum())
        assert [f.index for f in kept] == np.flatnonzero(ref_keep).tolist()
        if ref_keep.any():
            best = min(kept, key=attrgetter("ctot"))
            assert best.index == int(g["argmin_kept"])
            n = int(g["nsteps"][best.index])
            assert_close(np.array(best.x), g["x"][best.index, :n], 1e-12, "best.x", atol_scale=1.0)
        else:
            with pytest.raises(ValueError):
                min(kept, key=attrgetter("ctot"))
        # off-road masks alone, straight from the device buffers
        eng = pl._engine
        eng.offroad(rw, None)
        right = eng.road_ok[0, :C].cpu().numpy() == 0
        eng.offroad(None, lw)
        left = eng.road_ok[0, :C].cpu().numpy() == 0
        assert (right == g["offroad_right"].astype(bool)).all() and (left == g["offroad_left"].astype(bool)).all()
        # the unfiltered optimum in Cartesian coordinates (frenet_to_glob of the Duckietown driver)
        assert pl.paths.index(pl.opt_path_tot) == int(g["argmin"])
        assert_close(dl.frenet_to_glob(lane, pl, proj), g["winner_xy"], 1e-12, "winner_xy", atol_scale=1.0)
        # same filter on plain lists (any list of Frenet-like objects, like the reference accepts)
        plain = [paths[c] for c in range(C)]
        for f in plain:
            f.x, f.y = [], []
        dl.frenet_to_glob_planner(pl, lane, plain, proj)
        kept2 = dl.check_paths(plain, obstacles, np.array([0.07, 0.0, 0.0]), IdentityMapper(), rw, lw)
        assert [f.index for f in kept2] == np.flatnonzero(ref_keep).tolist()


def test_duckietown_lane_reference_style_trajectory_and_step():
    """A bare object carrying the mapper's lambdas (lib/mapper/mapper_obstacles.py:169-186) is accepted, and lane_step chains
    replan -> transform -> filter -> pick like the reference's per-frame loop (:222-238)."""
    from otg_b200.drivers import dt_lane as dl

    class Bare:
        pass

    a, b, c = 0.3, 0.05, 0.01
    tr = Bare()
    tr.compute_pt = lambda x: np.array([x, a * x ** 2 + b * x + c])
    tr.compute_first_derivative = lambda x: np.array([1, 2 * a * x + b])
    tr.compute_second_derivative = lambda x: np.array([0, 2 * a])
    lane = dl.lane_trajectory(tr)
    assert (lane.a, lane.b, lane.c) == (a, b, c)

    class IdentityMapper:
        def cam2rob(self, pts):
            return np.asarray(pts, dtype=np.float64)

    K, Pm = _classes()["dt_obstacles"]
    pl = K(Pm())
    pl.initialize(t0=0, p0=(0.01, 0.0, 0.0), s0=(0.0, 0.2, 0.0))
    obstacles = [{"end_point": (0.6, 0.05), "center": (0.6, 0.05), "end_right": (0.6, 0.05), "end_lat": (0.58, 0.12), "end_top": (0.58, 0.12)}]
    pos_s, pos_d, kept, xy = dl.lane_step(pl, tr, 0.07, 1 / 30, obstacles, IdentityMapper(), np.array([0.32, 0.05, -0.29]),
                                          np.array([0.28, 0.05, 0.33]))
    assert 0 < len(kept) < len(pl.paths)
    assert xy.shape == (len(pl.opt_path_tot.t), 2)
    # the pick is collision free and on the road: oracle check on the winner's own samples
    x, y = xy[:, 0][None, :], xy[:, 1][None, :]
    n = np.array([xy.shape[0]])
    pts = np.array([(0.6, 0.05), (0.58, 0.12)])
    assert fo.check_collisions(x, y, n, pts, fo.AGENT_SAFETY_RAD ** 2)[0][0]
    assert not fo.offroad(x, y, n, (0.32, 0.05, -0.29), "right")[0] and not fo.offroad(x, y, n, (0.28, 0.05, 0.33), "left")[0]
    assert pl.opt_path_tot.ctot == min(f.ctot for f in kept)


@pytest.mark.parametrize("lattice", ["default", "dense"])
def test_speculative_replan_equals_the_plain_call_sequence(lattice):
    """A planner remembers the path frenet_to_glob was given and - once check_paths has seen the same obstacle set twice - the
    obstacles, and its next replan runs transform, collision stage and the masked pick in the same launch (one kernel:
    FRENET_STAGE_ONE_LAUNCH); frenet_to_glob / check_paths then only validate.  Every call of the sequence must return what a
    planner without any of this returns (separate launches, device-side I/O blocks, no speculation), on a static scene (hits),
    after the scene changes (a miss, then hits again) and when nothing is sensed."""
    import otg_b200  # noqa: F401
    from otg_b200 import drivers, workloads as w
    from otg_b200.drivers import moving_obstacle_planner as mv
    from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
    from otg_b200.sensor import StaticObstacle
    from otg_b200.trajectory import CircleTrajectory2D, SplineTrajectory2D
    if lattice == "dense":
        sc = w.config4_scenario()
        tr = SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY)
        prm = dict(w.DENSE_PARAMS)
        init = dict(t0=sc["t0"], p0=sc["p0"], s0=sc["s0"], dd=sc["dd"])
        xy = np.stack(fo.frenet_to_cartesian(fo.SplinePath(w.SPLINE_WX, w.SPLINE_WY), sc["obs_s"], sc["obs_d"]), 1)[:40]
    else:
        tr = CircleTrajectory2D(8, 5, 2)
        prm = {}
        init = dict(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(3.0, 2.0, 0.0))
        xy = np.array([[5.0, 0.3], [6.0, -1.0], [4.5, 2.5], [9.0, 1.0]])

    class Sees:
        def __init__(self):
            self.obstacles = [StaticObstacle(p.copy()) for p in xy]

        def sense(self, *a, **k):
            return list(self.obstacles)

    planners = [TrajectoryPlannerV1(TrajectoryPlannerParams(**prm), _n_obs_cap=64),
                TrajectoryPlannerV1(TrajectoryPlannerParams(**prm), _n_obs_cap=64, _speculate=False, _one_launch=False, _zero_copy=False),
                # inputs AND outputs in mapped pinned memory, one launch replayed from a graph (no inline argument)
                TrajectoryPlannerV1(TrajectoryPlannerParams(**prm), _n_obs_cap=64, _zero_copy=True),
                # the inputs as the kernel's argument (frenet_replan_inline) instead of a copy node
                TrajectoryPlannerV1(TrajectoryPlannerParams(**prm), _n_obs_cap=64, _inline=True)]
    sensors = [Sees(), Sees(), Sees(), Sees()]
    for pl in planners:
        pl.initialize(**init)
    hits = 0
    for step in range(12):
        t = init["t0"] + 0.1 * (step + 1)
        if step == 6:                      # the scene changes: one obstacle moves, one disappears
            for s in sensors:
                s.obstacles[0].position = s.obstacles[0].position + np.array([0.4, -0.2])
                s.obstacles.pop()
        if step == 10:                     # nothing in sight
            for s in sensors:
                s.obstacles = []
        out = []
        for pl, s in zip(planners, sensors):
            pos = pl.replanner(t)
            if pl is planners[0] and getattr(pl._engine, "spec_pos", None) is not None:
                hits += 1
            paths = drivers.frenet_to_glob(pl, tr, pl.paths)
            kept, leading = mv.check_paths(paths, s, None)
            pl.paths = kept
            best = drivers.best_path(kept)
            pl.opt_path_tot = best
            out.append((pos, len(kept), len(leading), best, [leading[i].position.tolist() for i in range(min(3, len(leading)))]))
        (pb, nb, lb, fb, bb) = out[1]
        for which, (pa, na, la, fa, ba) in enumerate((out[0], out[2], out[3])):
            assert pa == pb and na == nb and la == lb and ba == bb, (step, which)
            assert fa.index == fb.index and (fa.cd, fa.cv, fa.ct, fa.ctot) == (fb.cd, fb.cv, fb.ct, fb.ctot), (step, which)
            for name in NAMES9 + ("x", "y"):
                assert getattr(fa, name) == getattr(fb, name), (step, which, name)
        if step in (3, 8):                 # and element access through the lazy views
            assert [f.index for f in planners[0].paths[:5]] == [f.index for f in planners[1].paths[:5]]
    assert hits >= 5          # replans 3..6 and 9.. ran with the obstacles on board


def test_c_host_example_matches_the_python_planner(tmp_path):
    """examples/replan_c_host.c: a host without Python or PyTorch - buffers sized by frenet_workspace_bytes, one replan through
    frenet_replan_inline, results read from mapped host memory.  Built with nvcc against the in-tree library and compared with the
    drop-in planner on the same inputs: candidate count, the pick, its cost (bit for bit) and the state the next replan starts from."""
    import os
    import shutil
    import subprocess
    import otg_b200  # noqa: F401
    from otg_b200 import _native as nat
    from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(nat.LIB_PATH)
    exe = str(tmp_path / "replan_c_host")
    subprocess.run([nvcc, "-o", exe, os.path.join(root, "examples", "replan_c_host.c"), "-I", os.path.join(root, "include"), "-L", libdir,
                    "-lfrenet_b200", "-Xlinker", "-rpath", "-Xlinker", libdir], check=True, capture_output=True)
    for p0, s0, t0 in (((0.2, 0.1, 0.0), (3.0, 2.2, 0.1), 0.1), ((-0.7, 0.0, 0.05), (11.0, 0.6, -0.1), 2.0)):       # high- and low-speed rows
        out = subprocess.run([exe] + [repr(float(v)) for v in (*p0, *s0, t0)], check=True, capture_output=True, text=True).stdout.split("\n")
        head = dict(zip(out[0].split()[0::2], out[0].split()[1::2]))
        nxt = [float(v) for v in out[1].split()[1:]]
        pl = TrajectoryPlannerV1(TrajectoryPlannerParams())
        pl.initialize(t0=t0, p0=p0, s0=s0)
        best = pl.opt_path_tot
        assert int(head["abi"]) == nat.ABI_VERSION
        assert int(head["n_valid"]) == len(pl.paths) and int(head["argmin_ctot"]) == best.index and int(head["steps"]) == len(best.t)
        assert float(head["min_ctot"]) == best.ctot
        assert nxt == [best.d[1], best.dot_d[1], best.ddot_d[1], best.s[1], best.dot_s[1], best.ddot_s[1]]
