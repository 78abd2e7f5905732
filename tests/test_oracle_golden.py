"""Pins the CPU oracle (oracle/frenet_oracle.py) against outputs of the UNMODIFIED reference.

The reference ships no golden vectors for this path, so the fixtures under
tests/golden/ were produced by importing /root/reference in the build container
(tests/golden/make_golden.py).  Every oracle function used by the GPU parity tests is checked here.
"""
import numpy as np
import pytest

from _golden import (golden, prefixed, pack_lattice, compare_packed, assert_close, near_tie, ORACLE_RTOL)
from oracle import frenet_oracle as fo

CIRCLE = fo.CirclePath(8, 5, 2)


def spline():
    z = golden("paths")
    return fo.SplinePath(list(z["spline_wx"]), list(z["spline_wy"]))


# ------------------------------------------------------------------------------------------------
def test_kat_variants():
    z = golden("kat_variants")
    for i in range(int(z["n_cases"])):
        c = prefixed(z, f"c{i}_")
        variant = str(c["variant"])
        prm = fo.OracleParams.defaults(variant)
        L = fo.generate(prm, c["p0"], c["s0"], float(c["t0"]), dd=float(c["dd"]))
        assert tuple(prm.di_interval()) == tuple(c["di_interval"])
        assert np.allclose(prm.t_interval(), c["t_interval"], rtol=0, atol=0)
        got = pack_lattice(L)
        compare_packed(got, c, ORACLE_RTOL, f"kat{i}/{variant}")
        kept = L.kept_indices()
        for key in ("cd", "cv", "ctot"):
            am = fo.first_argmin(getattr(L, key), L.keep)
            assert int(np.searchsorted(kept, am)) == int(c[f"argmin_{key}"]), (i, key)
        assert L.feasible[L.keep].all()


def test_known_answers_from_survey():
    """SURVEY.md section 4 table (regenerated with the reference, NumPy 2.3.5)."""
    L = fo.generate(fo.OracleParams.defaults("v1"), (0, 0, 0), (0, 0, 0), 0.1)
    assert L.n_candidates == 320 and int(L.nsteps.sum()) == 5920
    assert sorted(set(L.nsteps.tolist())) == [11, 16, 21, 26]
    assert fo.first_argmin(L.ctot) == 312
    assert L.ctot[312] == pytest.approx(0.33450470400000004, rel=1e-12)
    L = fo.generate(fo.OracleParams.defaults("dt"), (0, 0, 0), (0, 0, 0), 0.1)
    assert L.n_candidates == 528 and int(L.nsteps.sum()) == 8928 and fo.first_argmin(L.ctot) == 316
    assert 12 not in set(L.nsteps.tolist())     # arange overshoot
    L = fo.generate(fo.OracleParams.defaults("dt_obstacles"), (0, 0, 0), (0, 0, 0), 0.1)
    assert L.n_candidates == 165 and int(L.nsteps.sum()) == 8415 and fo.first_argmin(L.ctot) == 71
    L = fo.generate(fo.OracleParams.defaults("optimal"), (3, .3, 0), (0, 2, -.5), 0.0)
    assert L.n_candidates == 378 and int(L.nsteps.sum()) == 11466 and fo.first_argmin(L.ctot) == 211
    assert L.ctot[211] == pytest.approx(0.14631858545806598, rel=1e-12)


# ------------------------------------------------------------------------------------------------
def test_paths_circle_and_spline():
    z = golden("paths")
    for name, path in (("circle", CIRCLE), ("spline", spline())):
        s, d = z[f"{name}_s"], z[f"{name}_d"]
        px, py = path.pt(s)
        assert_close(np.stack([px, py], 1), z[f"{name}_pt"], 1e-12, f"{name}.pt")
        gx, gy = path.tangent(s)
        assert_close(np.stack([gx, gy], 1), z[f"{name}_grad"], 1e-9, f"{name}.grad", atol_scale=60.0)
        x, y = fo.frenet_to_cartesian(path, s, d)
        assert_close(np.stack([x, y], 1), z[f"{name}_xy"], 1e-12, f"{name}.xy", atol_scale=60.0)
    sp = spline()
    assert_close(sp.knots, z["spline_knots"], 1e-15, "knots")
    for j, k in enumerate("abcd"):
        n = len(z[f"spline_{k}x"])
        assert_close(np.asarray(sp.cx[j])[:n], z[f"spline_{k}x"], 1e-12, f"spline.{k}x", atol_scale=10.0)
        assert_close(np.asarray(sp.cy[j])[:n], z[f"spline_{k}y"], 1e-12, f"spline.{k}y", atol_scale=10.0)


# ------------------------------------------------------------------------------------------------
V1 = fo.OracleParams.defaults("v1")


def _replay(z, kind):
    """Open-loop replay of every recorded replan of a closed-loop reference run."""
    n = len(z["time"])
    cost_steps = {int(s): j for j, s in enumerate(z["cost_steps"])}
    full_steps = set(int(s) for s in z["full_steps"])
    prev = None     # (Lattice, winner index) of the previous replan
    n_masked_diff = 0
    for i in range(n):
        t_i = float(z["time"][i])
        if prev is None:
            Lp = fo.generate(V1, z["init_p0"], z["init_s0"], t_i)
            compare_packed(pack_lattice(Lp), prefixed(z, "init_"), ORACLE_RTOL, f"{kind}/init")
            assert fo.first_argmin(Lp.ctot) == int(z["init_argmin"])
            prev = (Lp, int(z["init_argmin"]))
        # closed-loop chain: new state = previous winner at time t_i (trajectory_planner.py:206-212)
        Lp, w = prev
        r, N = Lp.row_of[w], Lp.nsteps[w]
        tw = Lp.t[r, :N]
        p0 = fo.optimal_at_time(t_i, tw, Lp.lat[w, 0, :N], Lp.lat[w, 1, :N], Lp.lat[w, 2, :N], V1.delta_t)
        s0 = fo.optimal_at_time(t_i, tw, Lp.lon[r, 0, :N], Lp.lon[r, 1, :N], Lp.lon[r, 2, :N], V1.delta_t)
        assert_close(np.array(p0), z["p0"][i], 1e-9, f"{kind}/step{i}.p0", atol_scale=1.0)
        assert_close(np.array(s0), z["s0"][i], 1e-9, f"{kind}/step{i}.s0", atol_scale=1.0)
        # open loop from the recorded state (exactly what the reference generated from)
        L = fo.generate(V1, z["p0"][i], z["s0"][i], t_i)
        assert L.n_candidates == int(z["ncand"][i])
        dsi = z["dsi"][i]
        assert L.V[0] == dsi[0] and len(L.V) == len(np.arange(*dsi))
        am = fo.first_argmin(L.ctot)
        assert am == int(z["argmin"][i]), f"{kind}/step{i}: argmin {am} vs {int(z['argmin'][i])}"
        assert L.ctot[am] == pytest.approx(float(z["ctot_min"][i]), rel=1e-9)
        C = L.n_candidates
        if i in cost_steps:
            j = cost_steps[i]
            for k in ("cd", "cv", "ctot"):
                assert_close(getattr(L, k), z[k][j, :C], ORACLE_RTOL, f"{kind}/step{i}.{k}")
        winner = am
        if kind != "full":
            ns = int(z["n_sensed"][i])
            obs = z["obs_xy"][i, :ns]
            x, y = fo.lattice_to_cartesian(CIRCLE, L)
            ok, first = fo.check_collisions(x, y, L.nsteps, obs)
            assert np.array_equal(ok.astype(np.uint8), z["keep"][i, :C]), f"{kind}/step{i}: collision mask"
            if kind == "moving":
                assert np.array_equal(first.astype(np.int8), z["first_blocker"][i, :C]), f"{kind}/step{i}: blockers"
            winner = fo.first_argmin(L.ctot, ok)
            assert winner == int(z["argmin_masked"][i]), f"{kind}/step{i}: masked argmin"
            n_masked_diff += int(winner != am)
            if i in full_steps:
                g = prefixed(z, f"s{i}_")
                nmax = g["x"].shape[1]
                valid = np.arange(nmax)[None, :] < L.nsteps[:, None]
                assert_close(np.where(valid, x[:, :nmax], np.nan), g["x"], 1e-12, f"{kind}/step{i}.x", atol_scale=20.0)
                assert_close(np.where(valid, y[:, :nmax], np.nan), g["y"], 1e-12, f"{kind}/step{i}.y", atol_scale=20.0)
        if i in full_steps:
            compare_packed(pack_lattice(L), prefixed(z, f"s{i}_"), ORACLE_RTOL, f"{kind}/step{i}")
        prev = (L, winner)
    return n_masked_diff


def test_sim_planner_full_replay():
    _replay(golden("sim_planner_full"), "full")


def test_sim_planner_static_replay():
    assert _replay(golden("sim_planner_static"), "static") == 86


def test_sim_planner_moving_replay():
    z = golden("sim_planner_moving")
    assert _replay(z, "moving") == 97
    assert int(z["follow_fired"].sum()) == 0


def test_moving_obstacle_positions():
    """lib/sensor/dynamic_obstacle.py:25-28, 43-56 through sensor.step_obstacle (proximity_sensor.py:82-89)."""
    z = golden("sim_planner_moving")
    for i in (0, 1, 17, 250, 498):
        x, y = fo.moving_obstacle_position(CIRCLE, z["time"][i], z["obs_offset"], z["obs_speed"], z["obs_time_offset"])
        assert_close(np.stack([x, y], 1), z["obs_all_xy"][i], 1e-12, f"moving obstacles @ step {i}", atol_scale=20.0)


def test_sensor_gate():
    z = golden("sensor_gate")
    for pose, obs, mask in zip(z["pose"], z["obs_xy"], z["mask"]):
        got = fo.sensor_gate(pose, obs, float(z["range"]), float(z["aperture"]))
        assert np.array_equal(got.astype(np.uint8), mask)
    for kind in ("static", "moving"):       # the closed-loop runs: which obstacles were sensed, in list order
        s = golden(f"sim_planner_{kind}")
        for i in range(0, 499, 7):
            got = fo.sensor_gate(s["robot_pose_in"][i], s["obs_all_xy"][i], 3.5, np.pi / 6)
            ns = int(s["n_sensed"][i])
            assert np.array_equal(np.nonzero(got)[0], s["obs_idx"][i, :ns]), (kind, i)


# ------------------------------------------------------------------------------------------------
def test_follow_mode_v1():
    z = golden("follow_mode")
    for j in range(3):
        g = prefixed(z, f"r{j}_")
        L = fo.generate(V1, g["p0"], g["s0"], float(g["time"]), target_triples=g["triples"])
        assert L.V[0] == g["si_interval"][0] and len(L.V) == len(np.arange(*g["si_interval"]))
        compare_packed(pack_lattice(L), g, ORACLE_RTOL, f"follow{j}")
        assert fo.first_argmin(L.ctot) == int(g["argmin"])


# OptimalParameters (lib/optimal_frenet_frame/utils.py:7-22)
OP = dict(p=(3, 0.3, 0), s=(0, 2, -0.5), s0_t=(1, 2, 0), s1_t=(8, 0, 0), Ts=10, t0=0, Tn=(0, 2.5, 5),
          s0_tb=(3, 2, 0), s1_tb=(10, 0, 0), px=(3, 0.2, 0), sx=(0.8, 2, -0.5))
TARGET_KEYS = ("t", "s", "dot_s", "ddot_s", "dddot_s")


def _oracle_targets():
    tg = fo.target_frenet(OP["s0_t"], OP["s1_t"], OP["Ts"])
    tgb = fo.target_frenet(OP["s0_tb"], OP["s1_tb"], OP["Ts"])
    return {"stop": tg, "merge": fo.merge_target(tg, tgb), "follow": fo.follow_target(tg, D0=2)}


def test_optimal_targets():
    """lib/optimal_frenet_frame/target.py:17-52 and the Optimal planner's follow mode
    (trajectory_planner_optimal.py:130-142), flows of lib/tests/test_optimal_frenet.py:60-102."""
    z = golden("optimal_targets")
    prm = fo.OracleParams.defaults("optimal")
    T = np.arange(*prm.t_interval())
    for name, tgt in _oracle_targets().items():
        for k in TARGET_KEYS:
            assert_close(tgt[k], z[f"{name}_target_{k}"], 1e-12, f"{name} target {k}", atol_scale=10.0)
        trip = fo.optimal_target_triples(tgt, T, OP["t0"], prm.delta_t)
        L = fo.generate(prm, OP["p"], OP["s"], OP["t0"], target_triples=trip)
        g = prefixed(z, f"{name}_init_")
        compare_packed(pack_lattice(L), g, ORACLE_RTOL, f"optimal/{name}/init")
        kept = L.kept_indices()
        assert int(np.searchsorted(kept, fo.first_argmin(L.ct, L.keep))) == int(g["argmin_ct"])
        assert int(np.searchsorted(kept, fo.first_argmin(L.ctot, L.keep))) == int(g["argmin_ctot"])
        for j, tn in enumerate(OP["Tn"]):        # replan_ct chain: recorded states -> same lattices
            g = prefixed(z, f"{name}_ct{j}_")
            trip = fo.optimal_target_triples(tgt, T, tn, prm.delta_t)
            L = fo.generate(prm, g["p0"], g["s0"], tn, target_triples=trip)
            compare_packed(pack_lattice(L, t_field=False), g, ORACLE_RTOL, f"optimal/{name}/ct{j}", fields=("s", "d"))
            kept = L.kept_indices()
            assert int(np.searchsorted(kept, fo.first_argmin(L.ct, L.keep))) == int(g["argmin_ct"])
    for j, tn in enumerate(OP["Tn"]):            # replan_cd_cv chain
        g = prefixed(z, f"cdcv{j}_")
        L = fo.generate(prm, g["p0"], g["s0"], tn)
        compare_packed(pack_lattice(L, t_field=False), g, ORACLE_RTOL, f"optimal/cdcv{j}", fields=())
        kept = L.kept_indices()
        assert int(np.searchsorted(kept, fo.first_argmin(L.cd, L.keep))) == int(g["argmin_cd"])
        assert int(np.searchsorted(kept, fo.first_argmin(L.cv, L.keep))) == int(g["argmin_cv"])
    # attribute-mutation flow (test_optimal_frenet.py:98-102): custom t/d intervals and klat
    g = prefixed(z, "mutated_")
    L = fo.generate(prm.with_(klat=0.3), g["p0"], g["s0"], 0.0, t_interval=(prm.min_t, prm.max_t, prm.dt / 2),
                    di_interval=(-1, 2.45, 0.2))
    compare_packed(pack_lattice(L), g, ORACLE_RTOL, "optimal/mutated")
    kept = L.kept_indices()
    assert int(np.searchsorted(kept, fo.first_argmin(L.ctot, L.keep))) == int(g["argmin"])


# ------------------------------------------------------------------------------------------------
# BASELINE configs 4 and 5: dense lattice on the spline path with obstacles (reference run offline)
NAMES11 = ("t", "d", "dot_d", "ddot_d", "dddot_d", "s", "dot_s", "ddot_s", "dddot_s", "x", "y")


def _dense_oracle(p0, s0, t0, dd, obs_xy):
    from otg_b200 import workloads
    prm = fo.OracleParams.defaults("v1").with_(**workloads.DENSE_PARAMS)
    sp = fo.SplinePath(workloads.SPLINE_WX, workloads.SPLINE_WY)
    L = fo.generate(prm, p0, s0, t0, dd=dd)
    x, y = fo.lattice_to_cartesian(sp, L)
    ok, first = fo.check_collisions(x, y, L.nsteps, obs_xy)
    return L, x, y, ok, first


def _field_arrays(L, x, y):
    """[11][C][Nmax] in NAMES11 order."""
    lat = [L.lat[:, j, :] for j in range(4)]
    lon = [L.lon[L.row_of, j, :] for j in range(4)]
    return [L.t[L.row_of, :]] + lat + lon + [x, y]


def test_dense_config4():
    import otg_b200  # noqa: F401
    from otg_b200 import workloads
    z = golden("dense_config4")
    sc = workloads.config4_scenario()
    sp = fo.SplinePath(workloads.SPLINE_WX, workloads.SPLINE_WY)
    ox, oy = fo.frenet_to_cartesian(sp, sc["obs_s"], sc["obs_d"])
    assert_close(np.stack([ox, oy], 1), z["obs_xy"], 1e-12, "config4 obstacles", atol_scale=60.0)
    L, x, y, ok, first = _dense_oracle(sc["p0"], sc["s0"], sc["t0"], sc["dd"], z["obs_xy"])
    C = L.n_candidates
    assert C == 9600 and int(L.nsteps.sum()) == 954720
    assert np.array_equal(L.nsteps, z["nsteps"])
    for k in ("cd", "cv", "ctot"):
        assert_close(getattr(L, k), z[k], ORACLE_RTOL, f"config4.{k}")
    assert fo.first_argmin(L.ctot) == int(z["argmin"]) == 6361
    assert L.ctot[6361] == pytest.approx(0.33889646121037686, rel=1e-12)
    assert np.array_equal(ok.astype(np.uint8), z["keep"])
    assert np.array_equal(first.astype(np.int16), z["first_blocker"])
    assert fo.first_argmin(L.ctot, ok) == int(z["argmin_masked"])
    arrs = _field_arrays(L, x, y)
    valid = np.arange(x.shape[1])[None, :] < L.nsteps[:, None]
    chk = np.stack([np.where(valid, a, 0.0).sum(axis=1) for a in arrs], 1)
    assert_close(chk, z["checksum"], 1e-9, "config4 checksums", atol_scale=float(np.abs(z["checksum"]).max()))
    last = np.stack([a[np.arange(C), L.nsteps - 1] for a in arrs], 1)
    assert_close(last, z["last"], 1e-9, "config4 last samples", atol_scale=60.0)
    idx = z["sample_idx"]
    for name, a in zip(NAMES11, arrs):
        assert_close(np.where(valid[idx], a[idx], np.nan), z[f"sample_{name}"], 1e-9, f"config4 sample {name}",
                     atol_scale=60.0)


def test_config5_sample():
    import otg_b200  # noqa: F401
    from otg_b200 import workloads
    z = golden("config5_sample")
    sc = workloads.config5_scenarios()
    sp = fo.SplinePath(workloads.SPLINE_WX, workloads.SPLINE_WY)
    for i, b in enumerate(z["scenario_idx"]):
        fs, fd = workloads.moving_obstacle_frenet(sc["obs_s"][b], sc["obs_d"][b], sc["obs_speed"][b],
                                                  sc["obs_offset"][b], sc["t0"][b])
        ox, oy = fo.frenet_to_cartesian(sp, fs, fd)
        assert_close(np.stack([ox, oy], 1), z["obs_xy"][i], 1e-12, f"config5[{b}] obstacles", atol_scale=60.0)
        L, x, y, ok, first = _dense_oracle(sc["p0"][b], sc["s0"][b], float(sc["t0"][b]), float(sc["dd"][b]), z["obs_xy"][i])
        C = L.n_candidates
        assert C == int(z["ncand"][i])
        assert np.array_equal(L.nsteps, z["nsteps"][i, :C])
        assert_close(L.cd, z["cd"][i, :C], ORACLE_RTOL, f"config5[{b}].cd")
        assert_close(L.ctot, z["ctot"][i, :C], ORACLE_RTOL, f"config5[{b}].ctot")
        assert fo.first_argmin(L.ctot) == int(z["argmin"][i])
        assert np.array_equal(ok.astype(np.uint8), z["keep"][i, :C])
        assert np.array_equal(first.astype(np.int16), z["first_blocker"][i, :C])
        assert fo.first_argmin(L.ctot, ok) == int(z["argmin_masked"][i])


# ------------------------------------------------------------------------------------------------
def test_dt_lane():
    """parabola path + projection-relative transform + off-road / two-set collision filter
    against the reference's Duckietown driver functions (lib/tests/test_mapper_planner_obstacles.py:25-135)."""
    z = golden("dt_lane")
    assert fo.AGENT_SAFETY_RAD == float(z["agent_safety_rad"])
    prm = fo.OracleParams.defaults(fo.VARIANT_DTOBS)
    for i in range(int(z["n_cases"])):
        c = prefixed(z, f"c{i}_")
        L = fo.generate(prm, c["p0"], c["s0"], 1 / 30)
        compare_packed(pack_lattice(L), c, ORACLE_RTOL, f"dt_lane{i}")
        path = fo.ParabolaPath(*c["fit"])
        nx, ny = path.tangent(np.array([0.0, 0.1, 0.37, -0.05]))
        th = np.arctan2(ny, nx)
        assert_close(np.stack([-np.sin(th), np.cos(th)], axis=1), c["normal"], 1e-15, "dt normal")
        kept = L.kept_indices()
        s = L.lon[L.row_of[kept], 0, :]
        x, y = fo.frenet_to_cartesian(path, np.where(np.isnan(s), 0.0, float(c["proj"]) + (s - float(c["s0"][0]))), L.lat[kept, 0, :])
        n = L.nsteps[kept]
        for j in range(len(kept)):
            assert_close(x[j, :n[j]], c["x"][j, :n[j]], 1e-13, "dt x", atol_scale=1.0)
            assert_close(y[j, :n[j]], c["y"][j, :n[j]], 1e-13, "dt y", atol_scale=1.0)
        keep = np.ones(len(kept), dtype=bool)
        for fit, side in ((c["rw"], "right"), (c["lw"], "left")):
            if np.isnan(fit).any():
                assert not c[f"offroad_{side}"].any()
                continue
            off = fo.offroad(x, y, n, fit, side)
            assert (off == c[f"offroad_{side}"].astype(bool)).all(), (i, side)
            keep &= ~off
        obs = np.concatenate([c["obs_r"], c["obs_lat"]])
        ok, _ = fo.check_collisions(x, y, n, obs, fo.AGENT_SAFETY_RAD ** 2)
        keep &= ok
        assert (keep == c["keep"].astype(bool)).all(), i
        am = fo.first_argmin(L.ctot[kept], keep) if keep.any() else -1
        assert am == int(c["argmin_kept"])
        w = int(c["argmin"])
        assert_close(np.stack([x[w, :n[w]], y[w, :n[w]]], axis=1), c["winner_xy"], 1e-13, "dt winner", atol_scale=1.0)
