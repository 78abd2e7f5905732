"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header declares, and the host-side
logic (lattice rules, layouts, parameter classes, sharding, host restatements of the per-step sim pieces)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from _golden import golden, prefixed, assert_close, ROOT
from oracle import frenet_oracle as fo

import otg_b200  # noqa: F401
from otg_b200 import _native as nat
from otg_b200 import lattice as lat
from otg_b200.engine import LatticeGeometry
from otg_b200 import planner as P

HEADER = os.path.join(ROOT, "include", "frenet_b200.h")


def test_library_exports_every_declared_symbol():
    text = open(HEADER).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(frenet_\w+)\s*\(", text, flags=re.M))
    assert len(declared) >= 14
    lib = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(nat.EXPORTS), "binding and header disagree"
    assert lib.frenet_abi_version() == int(re.search(r"FRENET_B200_ABI_VERSION (\d+)", text).group(1))
    out = subprocess.run(["cuobjdump", "-lelf", nat.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_the_header():
    """Field order of the ctypes mirrors == field order in the header (a silent mismatch would corrupt launches)."""
    text = open(HEADER).read()
    for cname, ctype in (("FrenetParams", nat.FrenetParams), ("FrenetLattice", nat.FrenetLattice), ("FrenetPath", nat.FrenetPath),
                         ("FrenetResult", nat.FrenetResult), ("FrenetBuffers", nat.FrenetBuffers), ("FrenetRoad", nat.FrenetRoad),
                         ("FrenetSizes", nat.FrenetSizes)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.search(r"(\w+)\s*(?:\[\d+\])?\s*$", part.strip()).group(1))
        assert names == [f[0] for f in ctype._fields_], cname
    assert ctypes.sizeof(nat.FrenetResult) == 64
    assert np.dtype(nat.RESULT_DTYPE).itemsize == 64


def test_workspace_sizes_and_launch_shape_from_the_abi_alone():
    """frenet_workspace_bytes / frenet_k2_launch_shape are host-only: a C host sizes every buffer without the Python engine
    (SURVEY section 8b `frenet_plan_workspace_bytes`).  Checked against the layout rules of the header for BASELINE config 5."""
    from otg_b200 import lattice as lat, workloads as w
    from otg_b200.engine import LatticeGeometry
    from otg_b200.planner import TrajectoryPlannerParams
    p = TrajectoryPlannerParams(**w.DENSE_PARAMS)
    for pad, f32 in ((4, 0), (8, 1)):
        g = LatticeGeometry.from_intervals(lat.di_interval(p, "v1"), lat.t_interval(p, "v1"), p.delta_t, pad)
        L = nat.FrenetLattice()
        L.n_scen, L.n_v_cap, L.n_t, L.n_d, L.n_pad_max, L.lat_f32 = 512, 7, g.n_t, g.n_d, g.n_pad_max, f32
        L.row_elems = 7 * g.sum_pad
        L.cand_elems = L.row_elems * g.n_d
        sz = nat.FrenetSizes()
        assert nat.lib.frenet_workspace_bytes(ctypes.byref(L), 128, 0, ctypes.byref(sz)) == nat.OK
        R, Cc = 7 * g.n_t, 7 * g.n_t * g.n_d
        assert (sz.rows, sz.candidates) == (512 * R, 512 * Cc)
        assert sz.lat == 512 * L.cand_elems * 6 * (4 if f32 else 8) and sz.lon == 512 * L.row_elems * 4 * 8
        assert sz.cost == 4 * 512 * Cc * 8 and sz.result == 512 * 64 and sz.winner == 512 * 12 * g.n_pad_max * 8
        assert sz.obs_xy == 512 * 128 * 16 and sz.obs_xy_t == 0 and sz.coef_lat == 512 * Cc * 48
        fields = [f[0] for f in nat.FrenetSizes._fields_][2:-1]
        assert sz.total == sum((getattr(sz, k) + 255) // 256 * 256 for k in fields)
        assert sz.total < 180e9
    rows, parts, grid, smem = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64()
    path = nat.FrenetPath()
    path.kind, path.n_knots = nat.PATH_SPLINE, 6
    assert nat.lib.frenet_k2_launch_shape(ctypes.byref(L), ctypes.byref(path), 128, 0, 1, ctypes.byref(rows), ctypes.byref(parts),
                                          ctypes.byref(grid), ctypes.byref(smem)) == nat.OK
    assert rows.value == 1 and parts.value == 1 and grid.value == 512 * 7 * g.n_t and 16 * 1024 < smem.value < 56 * 1024
    L.n_scen = 1                       # one planner: the rows' candidates are split over several CTAs
    assert nat.lib.frenet_k2_launch_shape(ctypes.byref(L), ctypes.byref(path), 64, 0, 1, None, ctypes.byref(parts), ctypes.byref(grid), None) == nat.OK
    assert parts.value > 1 and grid.value == parts.value * 7 * g.n_t
    L.row_elems += 1
    assert nat.lib.frenet_workspace_bytes(ctypes.byref(L), 128, 0, ctypes.byref(sz)) == nat.ERR_BAD_ARG


def test_calls_fail_cleanly_without_a_device_or_buffers():
    """No CPU fallback: a launch without valid buffers is a bad-argument error, never a computed result."""
    L, Pm, B = nat.FrenetLattice(), nat.FrenetParams(), nat.FrenetBuffers()
    assert nat.lib.frenet_k2_evaluate(ctypes.byref(L), ctypes.byref(Pm), ctypes.byref(B), None) == nat.ERR_BAD_ARG
    assert nat.lib.frenet_plan(ctypes.byref(L), ctypes.byref(Pm), None, ctypes.byref(B), nat.STAGE_K1, 0, 0, None) == nat.ERR_BAD_ARG
    assert b"lattice" in nat.lib.frenet_last_error()
    import torch
    if not torch.cuda.is_available():
        from otg_b200.engine import FrenetEngine
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            FrenetEngine(1)
        pl = P.TrajectoryPlannerV1(P.TrajectoryPlannerParams())
        with pytest.raises(RuntimeError):
            pl.initialize(t0=0.1, p0=(0, 0, 0), s0=(0, 0, 0))


def test_parameter_classes_match_reference_defaults():
    z = golden("kat_variants")
    ref = {"v1": (P.TrajectoryPlannerParams, 0), "dt": (P.TrajectoryPlannerParamsDT, 1),
           "dt_obstacles": (P.TrajectoryPlannerParamsDTObstacles, 2), "optimal": (P.TrajectoryPlannerParamsOptimal, 3)}
    for variant, (cls, case) in ref.items():
        p = cls()
        g = prefixed(z, f"c{case}_")
        assert tuple(lat.di_interval(p, variant)) == tuple(g["di_interval"])
        assert np.array_equal(np.array(lat.t_interval(p, variant), dtype=np.float64), g["t_interval"])
        assert tuple(lat.dsi_interval(p, float(g["s0"][1]))) == tuple(g["dsi_interval"])
        o = fo.OracleParams.defaults(variant)
        for k in ("kj", "kt", "kd", "ks", "kdots", "klat", "klong", "low_speed_threshold", "desired_speed", "num_sample", "d_d_s"):
            assert getattr(p, k) == getattr(o, k), (variant, k)
    assert P.TrajectoryPlannerParams(kj=5).kj == 5 and P.TrajectoryPlannerParams().s_target is None
    assert P.TrajectoryPlannerDefaultParamsDTObstacles.sampling_t == 0.05 and P.TrajectoryPlannerParamsOptimal().S_thresh == 2
    assert not hasattr(P.TrajectoryPlannerParamsDT(), "dt")          # the DT variants have no `dt`
    f = P.Frenet()
    assert f.t == [] and f.x == [] and f.ctot == 0.0 and f.ct == 0.0


def test_lattice_axes_and_rounding_rules():
    p = P.TrajectoryPlannerParams()
    # banker's rounding on both ends: ds0 = 0.5 -> (-2, 4), 1.5 -> (-2, 6), 2.5 -> (0, 6)
    assert lat.dsi_interval(p, 0.5)[:2] == (-2, 4) and lat.dsi_interval(p, 2.5)[:2] == (0, 6)
    for ds0 in (0.0, 0.5, 1.5, 2.5, 3.49, 3.5, -0.5):
        a = lat.long_axis(p, ds0, False)
        assert len(a) <= lat.max_long_axis_len(p)
        assert np.array_equal(a, fo.long_axis(fo.OracleParams.defaults("v1"), ds0, False))
    # the vectorised axis of the batch planner == np.arange per scenario
    rng = np.random.default_rng(0)
    for q, n in ((1, 2), (1, 3), (0.5, 2), (0.2, 2)):
        ds = np.concatenate([rng.uniform(-1, 6, 500), np.arange(-1, 6, 0.25)])
        lo = np.round(ds - q * n)
        hi = np.round(ds + q * n + q)
        nv = np.ceil((hi - lo) / q).astype(np.int32)
        pp = P.TrajectoryPlannerParams(d_d_s=q, num_sample=n)
        assert nv.max() <= lat.max_long_axis_len(pp)
        for i in range(len(ds)):
            ref = np.arange(round(ds[i] - q * n), round(ds[i] + q * n + q), q)
            assert nv[i] == len(ref) and np.array_equal(lo[i] + np.arange(nv[i]) * ((lo[i] + q) - lo[i]), ref)
    # step counts: np.arange(0, T + dt, dt) overshoots (no 12 in the DT variant)
    g = LatticeGeometry.from_intervals(lat.di_interval(P.TrajectoryPlannerParamsDT(), "dt"),
                                       lat.t_interval(P.TrajectoryPlannerParamsDT(), "dt"), 0.1)
    assert 12 not in g.n_steps.tolist() and g.n_steps.min() == 11 and g.n_d == 8
    assert (g.n_pad % 4 == 0).all() and (g.n_pad >= g.n_steps).all() and g.t_offset[-1] == g.n_pad.sum()
    # the power table is what the reference's `**` gives
    for k in (0, 1, 7, g.n_pad_max - 1):
        t = np.arange(0, 10, 0.1)[k] if k < 100 else k * 0.1
        assert g.t_pow[k].tolist() == [t ** 2, t ** 3, t ** 4, t ** 5]


def test_closed_form_coefficients_match_linear_solve():
    from otg_b200.trajectory import QuinticPolynomial, QuarticPolynomial, quintic_bvp, quartic_bvp
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.uniform(-3, 3, 6)
        T = rng.uniform(0.5, 6)
        q = QuinticPolynomial(*a, T)
        ref = np.array([q.a0, q.a1, q.a2, q.a3, q.a4, q.a5])
        got = np.array(quintic_bvp(*a, T))
        assert np.allclose(got, ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max())
        assert np.allclose(ref, fo.quintic_coeffs(*a, T)[0], rtol=0, atol=0)
        k = QuarticPolynomial(a[0], a[1], a[2], a[3], 0.0, T)
        got = np.array(quartic_bvp(a[0], a[1], a[2], a[3], 0.0, T))
        assert np.allclose(got, [k.a0, k.a1, k.a2, k.a3, k.a4], rtol=1e-10, atol=1e-13)
        t = rng.uniform(0, T)
        assert q.compute_pt(T) == pytest.approx(a[3], abs=1e-9) and q.compute_first_derivative(T) == pytest.approx(a[4], abs=1e-9)
        assert k.compute_first_derivative(T) == pytest.approx(a[3], abs=1e-9)


def test_host_paths_and_descriptors():
    from otg_b200.trajectory import CircleTrajectory2D, SplineTrajectory2D, path_descriptor_of
    z = golden("paths")
    circ = CircleTrajectory2D(8, 5, 2)
    spl = SplineTrajectory2D(list(z["spline_wx"]), list(z["spline_wy"]))
    for name, tr in (("circle", circ), ("spline", spl)):
        idx = np.r_[0:40, len(z[f"{name}_s"]) - 12:len(z[f"{name}_s"])]
        pt = np.array([tr.compute_pt(s) for s in z[f"{name}_s"][idx]])
        g = np.array([tr.compute_first_derivative(s) for s in z[f"{name}_s"][idx]])
        assert_close(pt, z[f"{name}_pt"][idx], 1e-12, f"{name}.pt", atol_scale=60.0)
        assert_close(g, z[f"{name}_grad"][idx], 1e-9, f"{name}.grad", atol_scale=60.0)
    kind, n, blk, tab = path_descriptor_of(spl)
    assert kind == 1 and n == 6 and tab.shape == (6, 9)
    assert_close(tab[:, 0], z["spline_knots"], 1e-15, "knots")
    assert_close(tab[:5, 2], z["spline_bx"], 1e-12, "bx", atol_scale=10.0)
    kind, n, blk, tab = path_descriptor_of(circ)
    assert kind == 0 and tab is None and blk[4] == pytest.approx(2 * 8 + 4 * np.pi + 10)
    assert circ.compute_curvature(9.0) == 0.5 and circ.compute_curvature(1.0) == 0.0 and circ.compute_curvature(-1.0) == 0.5

    class Other:
        pass
    with pytest.raises(TypeError):
        path_descriptor_of(Other())


def test_sensor_and_sim_pieces_against_recorded_reference_runs():
    """ProximitySensor gate, Unicycle / controller / GN projection: one step of the recorded closed loop reproduces the
    next recorded pose when driven with the recorded planner output (host-side restatements)."""
    from otg_b200.sensor import ProximitySensor, StaticObstacle, MovingObstacle
    from otg_b200.sim import FrenetGNTransformOld, FrenetIOLController, Unicycle
    from otg_b200.trajectory import CircleTrajectory2D
    from otg_b200.drivers.obstacle_planner import seeded_obstacles
    z = golden("sensor_gate")
    for pose, obs, mask in zip(z["pose"], z["obs_xy"], z["mask"]):
        robot = Unicycle()
        robot.set_initial_pose(pose.copy())
        sensor = ProximitySensor(float(z["range"]), float(z["aperture"]))
        sensor.attach(robot)
        lst = [StaticObstacle(p) for p in obs]
        sensor.set_obstacle(lst)
        assert [1 if o in sensor.sense(np.zeros(3)) else 0 for o in lst] == mask.tolist()     # positional pose is ignored
    s = golden("sim_planner_full")
    tr = CircleTrajectory2D(8, 5, 2)
    transformer, controller = FrenetGNTransformOld(0.5, 10), FrenetIOLController(.5, 0.0, 5.0, 0.0, 0.0)
    robot = Unicycle()
    robot.set_initial_pose(np.array([0.0, 0.0, 0.0]))
    robot_p = robot.p
    transformer.estimatePosition(tr, robot_p)
    for i in range(60):
        est = transformer.estimatePosition(tr, robot_p)
        f = transformer.transform(robot_p)
        pos_s, pos_d = s["ret_s"][i], s["ret_p"][i]
        u = controller.compute(f, np.array([0, pos_d[0]]) - f[0:2], np.array([pos_s[1], pos_d[1]]), tr.compute_curvature(est))
        robot_p, _ = robot.step(u, 0.1)
        assert_close(robot_p, s["robot_pose"][i], 1e-10, f"pose after step {i}", atol_scale=10.0)
    m = golden("sim_planner_moving")
    pts = seeded_obstacles(np.arange(0.1, 50, 0.1), tr)
    assert_close(pts.T, m["obs_start"], 1e-13, "seeded obstacle positions", atol_scale=20.0)
    obs = [MovingObstacle(pts[:, i], tr) for i in range(10)]
    assert [o.trajectory.offset for o in obs] == m["obs_offset"].tolist()
    for i in (0, 100, 498):
        for o in obs:
            o.step(m["time"][i])
        assert_close(np.array([o.position for o in obs]), m["obs_all_xy"][i], 1e-13, f"moving obstacles @ {i}", atol_scale=20.0)


def test_workload_generators_are_deterministic_and_sharded_consistently():
    from otg_b200 import workloads as w
    from otg_b200.batch import shard_range
    a, b = w.config5_scenarios(), w.config5_scenarios()
    assert all(np.array_equal(a[k], b[k]) for k in a) and a["p0"].shape == (4096, 3) and a["obs_s"].shape == (4096, 128)
    assert (a["obs_s"] >= a["s0"][:, :1] + w.OBSTACLE_CLEARANCE).all() and a["obs_s"].max() < 54.9
    for world in (1, 2, 3, 8):
        parts = [shard_range(4096, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == 4096 and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    assert shard_range(10, 3, 4) == (8, 10) and shard_range(2, 3, 4) == (2, 2)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package, bench's GPU arm aside, may import or name it."""
    pkg = os.path.join(ROOT, "optimal-trajectory-generation-in-the-duckietown-environment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "frenet_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) needs no GPU: one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "candidate-timesteps/s"
    # the unmodified reference (oracle/_ref, a build-time copy made by oracle/build_ref.py) when it is there, else the NumPy port
    have_ref = os.path.isdir(os.path.join(root, "oracle", "_ref", "lib", "planner"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_winner_frenet_lists_are_made_on_first_use():
    """The planner's picks are `WinnerFrenet` objects: a private copy of the winner block whose rows become plain lists the first
    time a field is read (the reference's `Frenet` holds lists).  Values, list semantics, assignment, copies and the direct look-up
    `optimal_at_time` uses must agree with an eagerly built `Frenet`."""
    import copy
    import pickle
    from otg_b200.planner.frenet import Frenet, WinnerFrenet, LIST_FIELDS
    rows = np.arange(11 * 7, dtype=np.float64).reshape(11, 7) / 3.0
    f = WinnerFrenet(rows.copy(), 9, [1.5, 2.5, 3.5, 7.5], 42)
    assert isinstance(f, Frenet) and (f.cd, f.cv, f.ct, f.ctot, f.index) == (1.5, 2.5, 3.5, 7.5, 42)
    assert f._value("s", 2) == rows[5, 2] and f._value("t", -1) == rows[0, -1] and isinstance(f._value("d", 0), float)
    assert f.state_at(rows[0, 3], 1.0 / 3.0) == ((rows[1, 3], rows[2, 3], rows[3, 3]), (rows[5, 3], rows[6, 3], rows[7, 3]))
    assert f.state_at(-1.0, 1.0 / 3.0)[1][0] == rows[5, 0] and f.state_at(1e9, 1.0 / 3.0)[0][0] == rows[1, -1]       # both clamps
    assert f.x == [] and f.y == []                                   # gathered before the transform ran: no Cartesian samples
    assert f.state_at(0.0, 1.0) is None                               # a list exists now: the caller takes the list path
    for i, name in enumerate(LIST_FIELDS[:9]):
        got = getattr(f, name)
        assert isinstance(got, list) and got == rows[i].tolist() and getattr(f, name) is got
    f.t.append(99.0)                                                 # a list like any other from now on ...
    assert f._value("t", -1) == 99.0                                 # ... and a list that exists wins over the block
    f.s = [1.0, 2.0]
    assert f._value("s", -1) == 2.0 and f.s == [1.0, 2.0]
    g = WinnerFrenet(rows.copy(), 11, [0.0, 0.0, 0.0, 0.0], 3)
    assert g.x == rows[9].tolist() and g.y == rows[10].tolist()
    h = copy.deepcopy(g)
    assert h.d == g.d and h.index == 3 and h._value("dot_s", 1) == rows[6, 1]
    k = pickle.loads(pickle.dumps(g))
    assert k.ddot_d == rows[3].tolist()
    with pytest.raises(AttributeError):
        g.no_such_field
    lazy = WinnerFrenet(rows.copy(), 9, [0.0] * 4, 0)
    lazy._xy_source = lambda name: [1.0, 2.0] if name == "x" else [3.0, 4.0]      # what frenet_to_glob leaves behind
    assert lazy.x == [1.0, 2.0] and lazy.y == [3.0, 4.0]


def test_c_host_example_compiles_against_the_header():
    """examples/replan_c_host.c (a host that uses nothing but the C ABI) must keep compiling against include/frenet_b200.h; the GPU
    suite builds, runs and compares it with the Python planner."""
    import shutil
    import tempfile
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    with tempfile.TemporaryDirectory() as tmp:
        res = subprocess.run([nvcc, "-c", os.path.join(ROOT, "examples", "replan_c_host.c"), "-I", os.path.join(ROOT, "include"), "-o",
                              os.path.join(tmp, "replan_c_host.o")], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
