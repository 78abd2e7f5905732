#!/usr/bin/env python
"""Benchmark of the Frenet lattice planner hot path (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: BASELINE.json config 5, the batched moving-obstacle sweep - dense lattice (~1.1e4 candidates x
51..147 steps), spline reference path, 128 moving obstacles per scenario.  A STEP is one pass of the whole
pipeline (obstacle placement, K1..K5, winner gather) over this rank's shard of 512 scenarios; at 8 GPUs one
step is the full 4096-scenario sweep ("weak" scaling: 512 scenarios per GPU).  Metric: candidate-timesteps/s,
every one of them sampled, costed, feasibility-checked, transformed to Cartesian and collision-checked.

  value : device-timed (CUDA events), inputs resident in HBM.
  e2e   : the same through the public API (`BatchPlanner.stage_inputs/launch/result`) with HOST buffers: H2D of
          states + obstacle parameters from pinned memory and D2H of result records + winner trajectories
          inside the timed region.
  roofline / kernels : each kernel timed alone with CUDA events over the same buffers.
  cpu_baseline : the reference's CPU path on the host cores, bounded samples of the same workload:
          (i) the UNMODIFIED reference (oracle/_ref, a build-time copy of its pure-Python `lib/` package, see
          oracle/build_ref.py) single-threaded and in one process per core, (ii) the NumPy oracle (oracle/frenet_oracle.py,
          a vectorised port) in one process per core.  The scenarios computed on the CPU are compared with the GPU's
          records of the same scenarios after the timed region (`parity_spot_check`).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCEN_PER_GPU = 512
N_OBS = 128
METRIC = "candidate-timesteps/sec (feasibility+collision-checked)"


FIELD_NOTE = {"spread": "obstacles uniform over a 25 m x 16 m field starting 3 m ahead of each robot (round-1 parity fixtures)",
              "literal": "SURVEY 8(d) verbatim: s_o ~ U(s0, s0 + 15), d_o ~ N(0, 2), no clearance"}
# Bounded sample of the dense lattice for the unmodified reference (pure Python, ~2e4 candidate-timesteps/s per core): the rows of
# the scenario's centre target speed (num_sample = 0) and every fourth horizon (dt = 0.4): 400 of its ~1e4 candidates, same D axis,
# same sampling period, same 128 obstacles.  Candidate-timesteps/s is per unit of work, so the figure extrapolates to the full lattice.
REF_SLICE = dict(num_sample=0, dt=0.4)


def workload(lo, hi, field="spread"):
    import otg_b200  # noqa: F401
    from otg_b200 import workloads as w
    sc = w.config5_scenarios(field=field)
    t0 = sc["t0"][lo:hi]
    fs, fd = w.moving_obstacle_frenet(sc["obs_s"][lo:hi], sc["obs_d"][lo:hi], sc["obs_speed"][lo:hi], sc["obs_offset"][lo:hi], t0[:, None])
    return dict(p0=sc["p0"][lo:hi], s0=sc["s0"][lo:hi], dd=sc["dd"][lo:hi], t0=t0, obs_s=fs, obs_d=fd)


def dense_params():
    import otg_b200  # noqa: F401
    from otg_b200 import workloads as w
    from otg_b200.planner import TrajectoryPlannerParams
    return TrajectoryPlannerParams(**w.DENSE_PARAMS)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_one(args):
    """One scenario of the step through the NumPy port; returns the work done and what the GPU record must agree with."""
    p0, s0, t0, dd, fs, fd, slice_kw = args
    from otg_b200 import workloads as w
    from oracle import frenet_oracle as fo
    prm = fo.OracleParams.defaults("v1").with_(**w.DENSE_PARAMS).with_(**slice_kw)
    sp = fo.SplinePath(w.SPLINE_WX, w.SPLINE_WY)
    ox, oy = fo.frenet_to_cartesian(sp, fs, fd)
    L = fo.generate(prm, p0, s0, t0, dd=dd)
    x, y = fo.lattice_to_cartesian(sp, L)
    ok, _ = fo.check_collisions(x, y, L.nsteps, np.stack([ox, oy], 1))
    w_ = fo.first_argmin(L.ctot, ok)
    return dict(cand_ts=int(L.nsteps.sum()), argmin_masked=int(w_), n_survivors=int(ok.sum()),
                min_masked_ctot=float(L.ctot[w_]) if w_ >= 0 else float("inf"))


def _ref_one(args):
    """The same scenario (its REF_SLICE rows) through the UNMODIFIED reference (oracle/_ref)."""
    p0, s0, t0, dd, fs, fd, slice_kw = args
    from otg_b200 import workloads as w
    from oracle import ref_runner as rr
    return rr.replan_pipeline({**w.DENSE_PARAMS, **slice_kw}, (w.SPLINE_WX, w.SPLINE_WY), p0, s0, t0, dd, (fs, fd))


def _ref_latency(_):
    from oracle import ref_runner as rr
    a, b = rr.default_lattice_replan(3)
    return a, b


def cpu_replan_latency(n_calls=30):
    """Metric 2 on the host: one replan of the reference's default lattice (320 candidates, 5920 cand-ts) through the CPU port,
    single thread - (a) generate + argmin (config 1), (b) + Frenet -> Cartesian + collision filter against 3 obstacles +
    masked pick (configs 2/3).  The unmodified reference measured in the survey container: 70.6 ms and ~122 ms."""
    from oracle import frenet_oracle as fo
    prm = fo.OracleParams.defaults("v1")
    path = fo.CirclePath(8, 5, 2)
    obs = np.array([[5.0, 0.3], [6.0, -1.0], [4.5, 2.5]])
    t1, t3 = [], []
    for i in range(n_calls + 3):
        s0 = (3.0 + 0.01 * i, 2.0, 0.0)
        t = time.perf_counter()
        L = fo.generate(prm, (0.0, 0.0, 0.0), s0, 0.2)
        fo.first_argmin(L.ctot, L.keep)
        ta = time.perf_counter()
        x, y = fo.lattice_to_cartesian(path, L)
        ok, _ = fo.check_collisions(x, y, L.nsteps, obs)
        fo.first_argmin(L.ctot, ok & L.keep)
        tb = time.perf_counter()
        t1.append(ta - t)
        t3.append(tb - t)
    med = lambda v: float(np.median(np.array(v[3:])) * 1e6)   # noqa: E731
    return {"config1_replanner_p50_us": med(t1), "config3_replan_transform_collide_pick_p50_us": med(t3), "cores": 1, "kind": "port",
            "calls": n_calls}


def _jobs(wl, n, slice_kw):
    return [(wl["p0"][i], wl["s0"][i], float(wl["t0"][i]), float(wl["dd"][i]), wl["obs_s"][i], wl["obs_d"][i], slice_kw) for i in range(n)]


def cpu_pass(pool, wl, n, fn=_cpu_one, slice_kw=None):
    """n scenarios through `fn` in the pool; returns (per-scenario results, candidate-timesteps, seconds)."""
    jobs = _jobs(wl, n, slice_kw or {})
    t = time.perf_counter()
    res = pool.map(fn, jobs, chunksize=1) if pool is not None else [fn(j) for j in jobs]
    dt = time.perf_counter() - t
    return res, sum(r["cand_ts"] for r in res), dt


def have_ref():
    from oracle import ref_runner as rr
    return rr.available()


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores, same config and metric.
    With oracle/_ref present this is the UNMODIFIED reference (kind "reference"), each step one REF_SLICE sample of one scenario
    per core; otherwise the NumPy port on one whole scenario per core (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = cores
    wl = workload(0, n)
    ref = have_ref()
    fn, slice_kw = (_ref_one, REF_SLICE) if ref else (_cpu_one, {})
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_pass(pool, wl, n, fn, slice_kw)
        tot_c, tot_t = 0, 0.0
        for _ in range(args.steps):
            _, c, t = cpu_pass(pool, wl, n, fn, slice_kw)
            tot_c += c
            tot_t += t
    v = tot_c / tot_t
    if ref:
        sample = (f"per step: {n} scenarios of the shard, one per worker process ({cores} processes), each the rows of its centre target "
                  f"speed and every 4th horizon (num_sample=0, dt=0.4: 400 of ~1e4 candidates) through the unmodified reference "
                  f"(initialize + frenet_to_glob + check_paths + min); {tot_c // args.steps} candidate-timesteps per step")
    else:
        sample = f"{n} of the {SCEN_PER_GPU} scenarios of one GPU shard per step, {cores} worker processes, NumPy port (oracle/_ref missing)"
    line = {"metric": METRIC, "value": v, "unit": "candidate-timesteps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": v, "unit": "candidate-timesteps/s", "cores": cores, "kind": "reference" if ref else "port", "sample": sample},
            "e2e": {"value": v, "unit": "candidate-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_dict(n_gpus, field="spread"):
    return {"workload": "BASELINE config 5: batched moving-obstacle sweep, dense V1 lattice (d_road_width=0.1, dt=0.1, "
                        "num_sample=3, delta_t=0.02; 9600-11200 candidates x 51..147 steps), spline path, 128 moving "
                        f"obstacles per scenario, R=0.7; obstacle field '{field}': {FIELD_NOTE[field]}",
            "obstacle_field": field,
            "scenarios_per_gpu": SCEN_PER_GPU, "scenarios_per_step": SCEN_PER_GPU * n_gpus, "n_obstacles": N_OBS,
            "parallelism": f"scenario-sharded x{n_gpus}, one gather of result records",
            "l2": "each step writes ~28 GB of lattice state per GPU (>> 126 MB L2): no L2 flush needed"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling of SM clock and throttle reasons every 20 ms.  Started BEFORE the warm-up (nvidia-smi needs a few hundred
    milliseconds to come up on an 8-GPU box - longer than a 20-step timed region); `mark()` brackets the timed region and `stop()`
    reports the samples inside it, or - if the region was too short to catch one - the samples since the start of the warm-up,
    which ran the same kernels (`window` says which)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def mark(self):
        """First call: the timed region starts; second call: it ended."""
        if self.t0 is None:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        import datetime
        self.proc.terminate()
        out = self.proc.communicate()[0]
        rows = []
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[4:8]]))
            except ValueError:
                continue
        t0, t1 = self.t0 or 0.0, self.t1 or time.time()
        inside = [r for r in rows if t0 - 0.01 <= r[0] <= t1 + 0.01]
        window = "timed region"
        if not inside:
            inside, window = rows, "warm-up + timed region (no sample fell inside the timed region)"
        reasons = set()
        for r in inside:
            for name, on in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if on:
                    reasons.add(name)
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[2] for r in inside) if inside else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def time_kernels(bp, reps):
    """Each kernel alone, CUDA events on the launching stream, over the buffers of the last step."""
    import ctypes as C
    import torch
    from otg_b200 import _native as nat
    e = bp.engine
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L, P, Bf, path = C.byref(e.L), C.byref(bp.params), C.byref(e.B), C.byref(bp.path.desc)
    calls = {
        "k1_coefficients": lambda: nat.lib.frenet_k1_coefficients(L, P, Bf, stream),
        "k2k3k4_fused": lambda: nat.lib.frenet_k2_fused(L, P, path, Bf, 1, stream),
        "k2k3_fused": lambda: nat.lib.frenet_k2_fused(L, P, path, Bf, 0, stream),
        "k2_evaluate": lambda: nat.lib.frenet_k2_evaluate(L, P, Bf, stream),
        "k3_cartesian": lambda: nat.lib.frenet_k3_cartesian(L, P, path, Bf, stream),
        "k4_collision": lambda: nat.lib.frenet_k4_collision(L, P, Bf, stream),
        "k5_argmin": lambda: nat.lib.frenet_k5_argmin(L, Bf, bp.use_mask, stream),
        "gather_winner": lambda: nat.lib.frenet_gather_winner(L, P, Bf, nat.SEL_MASKED, 1, stream),
        # what the pipeline launches: argmin and winner gather as one kernel
        "k5_argmin_plus_gather": lambda: nat.lib.frenet_plan(L, P, None, Bf, nat.STAGE_K5 | nat.STAGE_GATHER, bp.use_mask, nat.SEL_MASKED, stream),
    }
    out = {}
    for name, fn in calls.items():
        nat.check(fn())
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record()
            nat.check(fn())
            b.record()
        torch.cuda.synchronize()
        out[name] = float(np.mean([a.elapsed_time(b) for a, b in evs]))   # ms
    return out


def predicted_mode(wl, steps, dev):
    """EXTENSION leg: the same step with obstacles propagated over the horizon on the device (time-indexed table,
    step k tested against the predicted position at t_initial + k * period) instead of the reference's snapshot."""
    import torch
    import otg_b200  # noqa: F401
    from otg_b200 import workloads as w
    from otg_b200.batch import BatchPlanner
    from otg_b200.trajectory import SplineTrajectory2D
    n = len(wl["t0"])
    sc = w.config5_scenarios()
    bp = BatchPlanner(dense_params(), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), n, N_OBS, "v1", dev, predict_steps=150)
    kw = dict(p0=wl["p0"], s0=wl["s0"], t0=wl["t0"], dd=wl["dd"], obs_s=sc["obs_s"][:n], obs_d=sc["obs_d"][:n] + sc["obs_offset"][:n],
              obs_speed=sc["obs_speed"][:n])
    bp.stage_inputs(**kw)
    bp.engine.upload()
    for _ in range(3):
        bp.launch_resident()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        bp.launch_resident()
        bp.engine.download(sync=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    feasible = int((bp.engine.result["status"] == 0).sum())
    # the two launches that differ from the snapshot step, each alone
    import ctypes as C
    from otg_b200 import _native as nat
    e = bp.engine
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    parts = {"predict_obstacles": lambda: e.predict_obstacles(bp.path, bp.period),
             "bounds_plus_fused_kernel": lambda: nat.check(nat.lib.frenet_k2_fused(C.byref(e.L), C.byref(bp.params), C.byref(bp.path.desc),
                                                                                   C.byref(e.B), 1, stream))}
    kernels = {}
    for name, fn in parts.items():
        fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for x, y in evs:
            x.record()
            fn()
            y.record()
        torch.cuda.synchronize()
        kernels[name] = {"ms": float(np.mean([x.elapsed_time(y) for x, y in evs]))}
    out = {"value": bp.cand_timesteps() / (ms * 1e-3), "unit": "candidate-timesteps/s", "ms_per_step": ms, "predict_steps": 150,
           "gpu_launches_per_step": bp.launches_per_plan, "scenarios_with_feasible": feasible, "kernels": kernels}
    del bp
    torch.cuda.empty_cache()
    return out


def all_checks_mode(wl, steps, dev):
    """EXTENSION leg: the same step with FINITE kinematic and curvature limits (v_max, a_max, kappa_max), i.e. every feasibility
    check of the evaluate kernel switched on next to the collision test - all in the one pipeline kernel.  The reference has no
    such limits (the headline step runs with +inf, where every candidate is feasible by definition)."""
    import torch
    import otg_b200  # noqa: F401
    from otg_b200 import workloads as w
    from otg_b200.batch import BatchPlanner
    from otg_b200.planner import TrajectoryPlannerParams
    from otg_b200.trajectory import SplineTrajectory2D
    n = len(wl["t0"])
    limits = dict(v_max=6.3, a_max=2.6, kappa_max=1.2)
    bp = BatchPlanner(TrajectoryPlannerParams(**w.DENSE_PARAMS, **limits), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), n, N_OBS, "v1", dev)
    bp.stage_inputs(**wl)
    bp.engine.upload()
    for _ in range(3):
        bp.launch_resident()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        bp.launch_resident()
        bp.engine.download(sync=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    e = bp.engine
    C_ = float(e.result["n_valid"].sum())
    kin = float(((e.cand_flags & 2) != 0).sum().item())
    curv = float(e.curv_ok.sum().item())
    out = {"value": bp.cand_timesteps() / (ms * 1e-3), "unit": "candidate-timesteps/s", "ms_per_step": ms, "limits": limits,
           "gpu_launches_per_step": bp.launches_per_plan, "kinematic_ok_fraction": kin / max(C_, 1.0),
           "curvature_ok_fraction": curv / max(C_, 1.0), "survivor_fraction": float(e.result["n_survivors"].sum()) / max(C_, 1.0),
           "scenarios_with_feasible": int((e.result["status"] == 0).sum())}
    del bp
    torch.cuda.empty_cache()
    return out


def _event_ms(fn, reps):
    import torch
    fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs]))


def literal_field_legs(dev, lo, hi, steps):
    """SURVEY 8(d)'s obstacle fields taken literally, next to the headline's: config 5 with s_o ~ U(s0, s0 + 15), d_o ~ N(0, 2)
    (same scenarios, same lattice) and config 4 with default_rng(64), s_o ~ U(5, 30), d_o ~ N(0, 2) on one planner."""
    import ctypes as C
    import torch
    import otg_b200  # noqa: F401
    from otg_b200 import _native as nat, workloads as w
    from otg_b200.batch import BatchPlanner
    from otg_b200.trajectory import SplineTrajectory2D
    spl = SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY)
    out = {}
    wl = workload(lo, hi, field="literal")
    bp = BatchPlanner(dense_params(), spl, hi - lo, N_OBS, "v1", dev)
    bp.stage_inputs(**wl)
    bp.engine.upload()
    for _ in range(3):
        bp.launch_resident()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        bp.launch_resident()
        bp.engine.download(sync=False)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    e = bp.engine
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    fused = _event_ms(lambda: nat.check(nat.lib.frenet_k2_fused(C.byref(e.L), C.byref(bp.params), C.byref(bp.path.desc), C.byref(e.B), 1, stream)), 5)
    bp.launch_resident()
    e.download(sync=True)
    nvalid = float(e.result["n_valid"].sum())
    out["config5_literal"] = {"value": bp.cand_timesteps() / (ms * 1e-3), "unit": "candidate-timesteps/s", "ms_per_step": ms,
                              "k2k3k4_fused_ms": fused, "survivor_fraction": float(e.result["n_survivors"].sum()) / max(nvalid, 1.0),
                              "scenarios_with_feasible": int((e.result["status"] == 0).sum()), "scenarios": hi - lo,
                              "field": FIELD_NOTE["literal"]}
    del bp
    torch.cuda.empty_cache()
    # config 4, both fields: one planner, 64 obstacles, all sensed
    for field in ("literal", "spread"):
        sc = w.config4_scenario(field=field)
        bp = BatchPlanner(dense_params(), spl, 1, 64, "v1", dev)
        kw = dict(p0=np.array([sc["p0"]]), s0=np.array([sc["s0"]]), t0=np.array([sc["t0"]]), dd=np.array([sc["dd"]]),
                  obs_s=sc["obs_s"][None], obs_d=sc["obs_d"][None])
        r = bp.plan(**kw)
        ms4 = _event_ms(bp.launch_resident, 20)
        out[f"config4_{field}"] = {"ms_per_replan_device": ms4, "candidate_timesteps": r.cand_timesteps,
                                   "value": r.cand_timesteps / (ms4 * 1e-3), "n_candidates": int(r.records["n_valid"][0]),
                                   "n_survivors": int(r.records["n_survivors"][0]), "argmin": int(r.records["argmin_ctot"][0]),
                                   "argmin_masked": int(r.records["argmin_masked"][0]), "field": FIELD_NOTE[field]}
        del bp
    torch.cuda.empty_cache()
    return out


def fp32_fast_path(wl, steps, dev, peak, records_f64):
    """Separate `dtype: "f32"` leg (never the headline): the candidate blocks are STORED as floats (24 B instead of 48 B per
    candidate-timestep, north star tolerance 1e-3); arithmetic, costs, collision masks and argmin stay fp64, so the records must
    equal the fp64 path's."""
    import ctypes as C
    import torch
    import otg_b200  # noqa: F401
    from otg_b200 import _native as nat, workloads as w
    from otg_b200.batch import BatchPlanner
    from otg_b200.trajectory import SplineTrajectory2D
    n = len(wl["t0"])
    bp = BatchPlanner(dense_params(), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), n, N_OBS, "v1", dev, lat_f32=True)
    bp.stage_inputs(**wl)
    bp.engine.upload()
    for _ in range(3):
        bp.launch_resident()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        bp.launch_resident()
        bp.engine.download(sync=False)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    e = bp.engine
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L, P, Bf, path = C.byref(e.L), C.byref(bp.params), C.byref(e.B), C.byref(bp.path.desc)
    fused = _event_ms(lambda: nat.check(nat.lib.frenet_k2_fused(L, P, path, Bf, 1, stream)), 10)
    k2k3 = _event_ms(lambda: nat.check(nat.lib.frenet_k2_fused(L, P, path, Bf, 0, stream)), 10)
    bp.launch_resident()
    e.download(sync=True)
    g = bp.geom
    cts = bp.cand_timesteps()
    nrow_ts, ncand = int(g.n_steps.astype(np.int64).sum()) * bp._n_v_sum, g.n_t * g.n_d * bp._n_v_sum
    by = 24 * cts + 32 * nrow_ts + ncand * (48 + 33)
    same = bool(records_f64 is not None and np.array_equal(np.asarray(e.result), np.asarray(records_f64)))
    out = {"dtype": "f32 storage of the candidate blocks (d and its three derivatives, x, y); fp64 arithmetic, costs, collision test, winner",
           "value": cts / (ms * 1e-3), "unit": "candidate-timesteps/s", "ms_per_step": ms,
           "k2k3k4_fused": {"ms": fused, "algorithmic_gb": (by + 5 * ncand) / 1e9, "gbs": (by + 5 * ncand) / (fused * 1e-3) / 1e9,
                            "frac_of_hbm_peak": (by + 5 * ncand) / (fused * 1e-3) / 1e9 / peak},
           "k2k3_fused": {"ms": k2k3, "algorithmic_gb": by / 1e9, "gbs": by / (k2k3 * 1e-3) / 1e9, "frac_of_hbm_peak": by / (k2k3 * 1e-3) / 1e9 / peak},
           "bytes_per_candidate_timestep": 24, "records_equal_fp64_path": same, "tolerance": "1e-3 relative on the stored samples (tests/test_gpu_parity.py)"}
    del bp
    torch.cuda.empty_cache()
    return out


def default_lattice_batch(dev, n_scen=32768, steps=10):
    """The reference's OWN lattice (configs 1-3: 320-384 candidates x 11..26 steps, circle path, 10 obstacles) as a batch of
    independent planners - the short-horizon form of the pipeline kernel."""
    import torch
    import otg_b200  # noqa: F401
    from otg_b200.batch import BatchPlanner
    from otg_b200.planner import TrajectoryPlannerParams
    from otg_b200.trajectory import CircleTrajectory2D
    rng = np.random.default_rng(1)
    B, O = n_scen, 10
    p0 = np.stack([rng.uniform(-1, 1, B), rng.uniform(-.3, .3, B), rng.uniform(-.3, .3, B)], 1)
    s0 = np.stack([rng.uniform(0, 30, B), rng.uniform(0, 4, B), rng.uniform(-.3, .3, B)], 1)
    bp = BatchPlanner(TrajectoryPlannerParams(), CircleTrajectory2D(8, 5, 2), B, O, "v1", dev)
    bp.pipeline_slots = 2          # 90 MB per output block here: two are enough on one GPU (no other rank to run ahead of)
    inputs = dict(p0=p0, s0=s0, t0=np.full(B, 0.1), obs_s=s0[:, 0:1] + rng.uniform(2, 12, (B, O)), obs_d=rng.uniform(-3, 3, (B, O)))
    bp.stage_inputs(**inputs)
    bp.engine.upload()
    for _ in range(3):
        bp.launch_resident()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        bp.launch_resident()
        bp.engine.download(sync=True)
    b.record()
    torch.cuda.synchronize()
    ms_serial = a.elapsed_time(b) / steps
    # the same with the 90 MB of winner trajectories leaving on a side stream under the next step's kernels (output blocks rotate)
    for _ in range(2):
        bp.launch_pipelined()
    bp.pipeline_wait()
    a.record()
    for _ in range(steps):
        bp.launch_pipelined()
    bp.pipeline_join()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    bp.pipeline_wait()
    a.record()
    for _ in range(steps):
        bp.launch_resident()
    b.record()
    torch.cuda.synchronize()
    ms_dev = a.elapsed_time(b) / steps
    cts = bp.cand_timesteps()
    out = {"planners": B, "ms_per_step": ms, "ms_per_step_download_on_the_compute_stream": ms_serial,
           "replans_per_s": B / (ms * 1e-3), "candidate_timesteps_per_s": cts / (ms * 1e-3),
           "ms_per_step_results_left_on_device": ms_dev, "candidate_timesteps_per_s_results_left_on_device": cts / (ms_dev * 1e-3),
           "d2h_bytes_per_step": bp.d2h_bytes, "candidate_timesteps_per_step": cts, "gpu_launches_per_step": bp.launches_per_plan,
           "planners_with_feasible": int((bp.engine.result["status"] == 0).sum())}
    del bp
    torch.cuda.empty_cache()
    return out


def closed_loop(dev, n_agents=8192, steps=30):
    """SURVEY section 8f row 1 leg: whole closed-loop steps (projection, replan, cone gate, collision filter, controller,
    unicycle) for `n_agents` agents of BASELINE config 2 (default lattice, circle path, 10 static obstacles) with all
    per-step state on the device.  The reference runs one agent at ~122 ms per step on one CPU thread."""
    import torch
    import otg_b200  # noqa: F401
    from otg_b200.batch_sim import BatchSimulator
    from otg_b200.drivers.obstacle_planner import seeded_obstacles
    from otg_b200.planner import TrajectoryPlannerParams
    from otg_b200.trajectory import CircleTrajectory2D
    tr = CircleTrajectory2D(8, 5, 2)
    pts = seeded_obstacles(np.arange(0.1, 50, 0.1), tr).T
    sim = BatchSimulator(TrajectoryPlannerParams(), tr, n_agents, 10)
    rng = np.random.default_rng(1)
    pose = np.stack([rng.uniform(0, 7, n_agents), rng.uniform(-0.3, 0.3, n_agents), rng.uniform(-0.1, 0.1, n_agents)], 1)
    sim.reset(pose, 0.1, obs_xy=np.repeat(pts[None], n_agents, 0))
    t_vect = np.arange(0.1, 50, 0.1)
    for i in range(5):
        sim.step(float(t_vect[i]))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(5, 5 + steps):
        sim.step(float(t_vect[i]))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    rec = sim.step(float(t_vect[5 + steps]), sync=True)
    cts = float(rec["n_valid"].sum()) / 64 * 16 * (11 + 16 + 21 + 26)
    out = {"agents": n_agents, "ms_per_step": ms, "agent_steps_per_s": n_agents / (ms * 1e-3), "candidate_timesteps_per_s": cts / (ms * 1e-3),
           "agents_with_feasible": int((rec["status"] == 0).sum()), "kernel_launches_per_step": 9}
    del sim
    torch.cuda.empty_cache()
    return out


def replan_latency(n_calls=200, **planner_kw):
    """Metric 2: wall time of ONE replan through the drop-in planner API, host perf_counter around
    the Python calls (H2D of inputs, kernels, D2H of the result record + winner trajectory, Python return).
    Reference figures measured in the survey container (1 CPU thread): 70.6 ms (config 1), ~122 ms (config 2), 14.3 s (dense)."""
    import otg_b200  # noqa: F401
    from otg_b200 import drivers, workloads as w
    from otg_b200.drivers import moving_obstacle_planner as mv
    from otg_b200.planner import TrajectoryPlannerV1, TrajectoryPlannerParams
    from otg_b200.sensor import ProximitySensor, StaticObstacle
    from otg_b200.sim import Unicycle
    from otg_b200.trajectory import CircleTrajectory2D, SplineTrajectory2D

    def pct(ts):
        ts = np.sort(np.array(ts)) * 1e6
        return {"p50_us": float(ts[len(ts) // 2]), "p90_us": float(ts[int(len(ts) * 0.9)]), "calls": len(ts)}

    out = {}
    # config 1: default lattice (320 candidates, 5920 cand-ts), planner.replanner only
    pl = TrajectoryPlannerV1(TrajectoryPlannerParams(), **planner_kw)
    pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(0.0, 0.0, 0.0))
    ts = []
    for i in range(n_calls + 20):
        t = time.perf_counter()
        pl.replanner(0.1 + 0.1 * (i + 1))
        ts.append(time.perf_counter() - t)
    out["config1_replanner"] = pct(ts[20:])
    # config 2/3: + frenet_to_glob + check_paths (3 sensed obstacles) + masked pick
    tr = CircleTrajectory2D(8, 5, 2)
    robot = Unicycle()
    robot.set_initial_pose(np.array([3.0, 0.0, 0.0]))
    sensor = ProximitySensor(3.5, np.pi / 6)
    sensor.attach(robot)
    sensor.set_obstacle([StaticObstacle(np.array(p)) for p in ([5.0, 0.3], [6.0, -1.0], [4.5, 2.5])])
    pl = TrajectoryPlannerV1(TrajectoryPlannerParams(), **planner_kw)
    pl.initialize(t0=0.1, p0=(0.0, 0.0, 0.0), s0=(3.0, 2.0, 0.0))
    ts = []
    for i in range(n_calls + 20):
        t = time.perf_counter()
        pl.replanner(0.2)
        paths = drivers.frenet_to_glob(pl, tr, pl.paths)
        pl.paths, _ = mv.check_paths(paths, sensor, robot.p)
        pl.opt_path_tot = drivers.best_path(pl.paths)
        ts.append(time.perf_counter() - t)
    out["config3_replan_transform_collide_pick"] = pct(ts[20:])
    # config 4: dense lattice (9600 candidates, 954720 cand-ts), spline path, 64 obstacles, one planner
    sc = w.config4_scenario()
    spl = SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY)
    pl = TrajectoryPlannerV1(TrajectoryPlannerParams(**w.DENSE_PARAMS), _n_obs_cap=64, **planner_kw)
    pl.initialize(t0=sc["t0"], p0=sc["p0"], s0=sc["s0"], dd=sc["dd"])
    eng = pl._engine
    ox, oy = eng.points_to_cartesian(eng.path_handle(spl), sc["obs_s"], sc["obs_d"])

    class AllSeeing:
        obstacles = [StaticObstacle(np.array([x, y])) for x, y in zip(ox, oy)]

        def sense(self, *a, **k):
            return self.obstacles

    ts = []
    for i in range(n_calls // 2 + 10):
        t = time.perf_counter()
        pl.replanner(sc["t0"])
        paths = drivers.frenet_to_glob(pl, spl, pl.paths)
        pl.paths, _ = mv.check_paths(paths, AllSeeing(), None)
        pl.opt_path_tot = drivers.best_path(pl.paths)
        ts.append(time.perf_counter() - t)
    out["config4_dense_64_obstacles"] = pct(ts[10:])
    return out


def spot_check(cpu_results, records):
    """CPU results (list of dicts from _cpu_one / _ref_one) against the GPU's FrenetResult records of the same scenarios:
    survivor counts exact, masked pick exact (a different index with a cost within 1e-6 relative counts as a tie, the north
    star's exception), minimum masked cost to 1e-9 relative."""
    n = len(cpu_results)
    mism, ties, worst = 0, 0, 0.0
    for r, g in zip(cpu_results, records):
        a, b = r["min_masked_ctot"], float(g["min_masked_ctot"])
        same_cost = (a == b) or (np.isfinite(a) and np.isfinite(b) and abs(a - b) <= 1e-9 * max(abs(a), abs(b)))
        if np.isfinite(a) and np.isfinite(b):
            worst = max(worst, abs(a - b) / max(abs(a), abs(b), 1e-300))
        if r["n_survivors"] != int(g["n_survivors"]) or not same_cost:
            mism += 1
        elif r["argmin_masked"] != int(g["argmin_masked"]):
            ties += 1       # equal minimum cost, different index: only possible for an exact / near tie
    return {"n": n, "mismatches": mism, "near_ties": ties, "max_rel_cost_diff": worst,
            "fields": ["n_survivors (exact)", "argmin_masked (exact, ties reported)", "min_masked_ctot (1e-9 rel)"]}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import otg_b200  # noqa: F401
    from otg_b200.batch import BatchPlanner, gather_results, shard_range
    from otg_b200.trajectory import SplineTrajectory2D
    from otg_b200 import workloads as w

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_total = SCEN_PER_GPU * world
    lo, hi = shard_range(n_total, rank, world)
    wl = workload(lo, hi)
    bp = BatchPlanner(dense_params(), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), hi - lo, N_OBS, "v1", dev)
    # N > 1, --balance-shards: the GPUs of one box differ by several percent on this kernel, and equal shards run at the slowest one's
    # pace.  A short calibration on the equal shards (every rank's own step time, no collective) sizes the shards by measured speed
    # (within +-10 % of the equal share; the total stays SCEN_PER_GPU x N scenarios per step).  Off by default: it only pays where the
    # spread is a stable property of the GPUs and larger than their run-to-run variation (measured both ways, profiles/r2_history.md).
    sizes, calib = None, None
    if world > 1 and args.balance_shards:
        from otg_b200.batch import balanced_shard_sizes, shard_range_from_sizes
        bp.stage_inputs(**wl)
        bp.engine.upload()
        for _ in range(3):
            bp.launch_pipelined(gather=False)
        bp.pipeline_wait()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        c0.record()
        for _ in range(6):
            bp.launch_pipelined(gather=False)
        bp.pipeline_join()
        c1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([c0.elapsed_time(c1) / 6], dtype=torch.float64, device=dev)
        allt = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt, mine)
        calib = [float(v) for v in allt.cpu().numpy()]
        sizes = balanced_shard_sizes(n_total, calib)
        if sizes[rank] != hi - lo or shard_range_from_sizes(sizes, rank) != (lo, hi):
            del bp
            torch.cuda.empty_cache()
            lo, hi = shard_range_from_sizes(sizes, rank)
            wl = workload(lo, hi)
            bp = BatchPlanner(dense_params(), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), hi - lo, N_OBS, "v1", dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(records=None, sync=True):
        # the one collective of the path: all ranks' 64-byte result records, all-gathered on the device
        return bp.gather_records(n_total, sync=sync, sizes=sizes) if world > 1 else (records if records is not None else bp.engine.result)

    # ---- resident arm: `value` ----
    bp.stage_inputs(**wl)
    bp.engine.upload()
    # N > 1: the timed steps are ONE sweep - every rank runs its steps at its own pace, every step's records stay on the device, and
    # the path's one collective gathers the records of all steps at the end of the sweep, inside the timed region
    # (BatchPlanner.sweep_begin / sweep_gather).  --gather-per-step runs the collective after every step on the side stream instead
    # (what the e2e arm below always does): on 4 and 8 GPUs the NCCL kernel spinning next to the pipeline kernel while it waits
    # for the slowest rank costs 1-2.5 % per step.
    per_step = world > 1 and (args.gather_per_step or sizes is not None)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        bp.launch_pipelined(n_total, sizes=sizes, gather=per_step)
    if world > 1 and not per_step:
        bp.sweep_begin(1)
        bp.launch_pipelined(n_total, gather=False)
        bp.sweep_gather(n_total)              # (warms the collective up)
    bp.pipeline_wait()
    barrier()
    if clocks:
        clocks.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    # Every step: the kernels on the compute stream; the D2H copy of its records + winners and (N > 1) the one collective of the
    # path - the all-gather of the records - on a side stream, under the next step's kernels (four output blocks rotate).  The
    # host does not wait per step (the e2e arm below is the one that hands every step's results to the host); the timed region
    # ends when the last step's results have arrived too.
    if world > 1 and not per_step:
        bp.sweep_begin(args.steps)
    for _ in range(args.steps):
        bp.launch_pipelined(n_total, sizes=sizes, gather=per_step)
    if world > 1 and not per_step:
        bp.sweep_gather(n_total, sync=False)
    bp.pipeline_join()
    ev1.record()
    barrier()
    bp.pipeline_wait()
    sweep_ok = None
    if world > 1 and not per_step:
        swr = bp.sweep_records()              # [steps][n_total]: the same inputs every step, so every step's records must agree
        sweep_ok = bool(all(np.array_equal(swr[0], swr[i]) for i in range(1, len(swr)))) and swr.shape == (args.steps, n_total)
    ms_dev = ev0.elapsed_time(ev1)
    if clocks:
        clocks.mark()
    clk = clocks.stop() if clocks else None
    # N > 1: the same loop WITHOUT the collective, so the line can say how much of the loss against N = 1 is the all-gather and
    # how much is the spread between the ranks' own GPUs (the job's time is the slowest rank's)
    ms_own = ms_dev
    if world > 1:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            bp.launch_pipelined(n_total, gather=False, sizes=sizes)
        bp.pipeline_join()
        ev1.record()
        barrier()
        bp.pipeline_wait()
        ms_own = ev0.elapsed_time(ev1)
    cts_step_local = bp.cand_timesteps()

    # ---- end-to-end arm: host buffers through the public API ----
    # Every step copies its own inputs host -> device and its results device -> host; the host stages the NEXT step's
    # inputs into the pinned block while the GPU works on the current one (the staging block is free again as soon as
    # the H2D copy has been consumed - `wait_staging_free`).
    # Three things overlap: the host stages step i + 1, the GPU runs step i, and step i - 1's results (D2H of records + winners and,
    # for N > 1, the all-gather of the records) travel on the side stream - `launch_streamed` / `collect`.  Every step's results are
    # read by the host inside the timed region, one step after they were produced.
    barrier()
    t0 = time.perf_counter()
    bp.stage_inputs(**wl)
    prev = None
    for i in range(args.steps):
        slot = bp.launch_streamed(n_total, sizes=sizes)
        if i + 1 < args.steps:
            bp.stage_inputs(**wl)
        if prev is not None:
            allrec = bp.collect(prev).records
        prev = slot
    allrec = np.array(bp.collect(prev).records, copy=True)
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0)

    t = torch.tensor([ms_dev, ms_e2e, float(cts_step_local)], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        mine = torch.tensor([ms_dev / args.steps, ms_own / args.steps], dtype=torch.float64, device=dev)
        allr = torch.empty(world * 2, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.cpu().numpy().reshape(world, 2)
        per_rank = {"collective": "after every step (side stream)" if per_step else "once per sweep of timed steps, inside the timed region",
                    "sweep_records_consistent": sweep_ok, "shard_sizes": sizes if sizes is not None else [SCEN_PER_GPU] * world, "calibration_ms_per_step_equal_shards": calib,
                    "ms_per_step_with_gather": [float(v) for v in allr[:, 0]], "ms_per_step_without_gather": [float(v) for v in allr[:, 1]],
                    "note": "device time of each rank's own timed loop; the job's ms_per_step is the maximum of the first list"}
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_dev, ms_e2e, cts_step = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        cts_step = float(cts_step_local)

    if rank == 0:
        kern = time_kernels(bp, max(3, args.steps))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        g = bp.geom
        cts, nrow_ts = cts_step_local, int(g.n_steps.astype(np.int64).sum()) * bp._n_v_sum
        ncand = g.n_t * g.n_d * bp._n_v_sum
        # algorithmic bytes per launch: what the kernel must read + write once
        percand = ncand * (48 + 33)            # K1's lateral coefficients in, 4 costs + flag out
        algo = {
            "k2k3k4_fused": 48 * cts + 32 * nrow_ts + percand + ncand * 5,
            "k2k3_fused": 48 * cts + 32 * nrow_ts + percand,
            "k2_evaluate": 32 * cts + 32 * nrow_ts + percand,
            "k3_cartesian": 24 * cts + 8 * nrow_ts,
            "k4_collision": 16 * cts + ncand * 5,
        }
        kernels = {}
        for name, ms in kern.items():
            kernels[name] = {"ms": ms}
            if name in algo:
                kernels[name].update(algorithmic_gb=algo[name] / 1e9, gbs=algo[name] / (ms * 1e-3) / 1e9,
                                     frac_of_hbm_peak=algo[name] / (ms * 1e-3) / 1e9 / peak)
        # the collision stage is read against the measured non-tensor FMA peak (scripts/alu_peak.cu, stored in
        # profiles/measured_alu_peaks.json).  A brute-force sweep costs 5 flop per (candidate-timestep, obstacle); the kernel's
        # cull skips most of it, so its brute-force-equivalent rate may exceed the fp64 peak.
        ap_path = os.path.join(ROOT, "profiles", "measured_alu_peaks.json")
        if os.path.exists(ap_path) and "k4_collision" in kernels:
            alu = json.load(open(ap_path))
            flops = 5.0 * N_OBS * cts
            eq = flops / (kern["k4_collision"] * 1e-3) / 1e12
            kernels["k4_collision"].update(brute_force_equiv_tflops=eq, fp64_fma_peak_tflops=alu["fp64_tflops"],
                                           fp32_fma_peak_tflops=alu["fp32_tflops"], equiv_over_fp64_peak=eq / alu["fp64_tflops"])
        in_pipeline = ("k1_coefficients", "k2k3k4_fused", "k5_argmin_plus_gather")   # what frenet_plan launches
        dom = max((k for k in algo if k in in_pipeline), key=lambda k: kern[k])
        pipeline_ms = sum(kern[k] for k in in_pipeline)
        # Bytes.  `frac` uses the bytes this design must move: the s-fields are stored once per ROW (the n_d = 80 candidates of a row
        # share them; the parity tests read all ten fields of every candidate through that view), i.e. 48 B per candidate-timestep
        # (d and its three derivatives, x, y) + 32 B per row-timestep.  SURVEY 8(d) counts the s-fields per candidate: 80 B per candidate-timestep
        # for the fused pipeline.  `frac_survey_bytes` is the same time against that figure; it can exceed 1 because those bytes
        # are never written.
        survey = {"k2k3k4_fused": 80 * cts, "k2k3_fused": 80 * cts, "k2_evaluate": 64 * cts + 40 * ncand}
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
                    "share_of_step": kern[dom] / pipeline_ms, "ms": kern[dom], "algorithmic_bytes": algo[dom],
                    "bytes_per_candidate_timestep": 48, "bytes_per_row_timestep": 32,
                    "frac_survey_bytes": survey[dom] / (kern[dom] * 1e-3) / 1e9 / peak, "survey_bytes_per_candidate_timestep": 80,
                    "bytes_note": "longitudinal samples stored once per row, not per candidate (row de-duplication): 48 B/cand-ts + 32 B/row-ts instead of SURVEY 8(d)'s 80 B/cand-ts"}
        tr = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tr):
            tj = json.load(open(tr))
            roofline["traffic"] = tj.get(dom)
            roofline["traffic_source"] = tj.get("source", "committed ncu capture (profiles/), not measured in this run")

        # CPU baseline on rank 0, N = 1 only: bounded samples of the same workload, and a spot check of the GPU's records
        cpu, spot = None, None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            n = min(2 * cores, SCEN_PER_GPU)
            spot = {}
            with mp.get_context("spawn").Pool(cores) as pool:
                cpu_pass(pool, wl, min(cores, n))   # warm the workers
                res_port, c, tcpu = cpu_pass(pool, wl, n)
                port = {"value": c / tcpu, "unit": "candidate-timesteps/s", "cores": cores, "kind": "port",
                        "sample": f"first {n} of the {SCEN_PER_GPU} scenarios of the step, NumPy port in {cores} processes, {tcpu:.1f} s"}
                spot["port"] = spot_check(res_port, allrec[:n])
                cpu = dict(port)
                cpu["port_all_cores"] = port
                cpu["replan_latency"] = cpu_replan_latency()
                if have_ref():
                    nref = min(cores, SCEN_PER_GPU)
                    pool.map(_ref_latency, [0] * cores, chunksize=1)      # warm the workers: import of the reference package
                    res_ref, c, tref = cpu_pass(pool, wl, nref, _ref_one, REF_SLICE)
                    what = ("rows of the centre target speed and every 4th horizon (num_sample=0, dt=0.4: 400 of ~1e4 candidates, same D axis, "
                            "sampling period and 128 obstacles) of one scenario, unmodified reference (initialize + frenet_to_glob + check_paths + min); "
                            "candidate-timesteps/s is per unit of work, so it extrapolates to the full lattice")
                    allc = {"value": c / tref, "unit": "candidate-timesteps/s", "cores": cores, "kind": "reference", "seconds": tref,
                            "sample": f"first {nref} scenarios of the step, one per process, each: {what}"}
                    _, c1, t1 = cpu_pass(None, wl, 1, _ref_one, REF_SLICE)
                    one = {"value": c1 / t1, "unit": "candidate-timesteps/s", "cores": 1, "kind": "reference", "seconds": t1,
                           "sample": f"first scenario of the step: {what}"}
                    lat_ref = pool.map(_ref_latency, [0])[0]
                    cpu.update(value=allc["value"], kind="reference", sample=allc["sample"], reference_all_cores=allc,
                               reference_single_thread=one)
                    cpu["replan_latency"]["reference_config1_replanner_p50_us"] = lat_ref[0] * 1e6
                    cpu["replan_latency"]["reference_config3_replan_transform_collide_pick_p50_us"] = lat_ref[1] * 1e6
                    # the GPU on the very same slice of the same scenarios
                    from otg_b200.planner import TrajectoryPlannerParams
                    bq = BatchPlanner(TrajectoryPlannerParams(**{**w.DENSE_PARAMS, **REF_SLICE}), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY),
                                      nref, N_OBS, "v1", dev)
                    rq = bq.plan(**{k: v[:nref] for k, v in wl.items()})
                    spot["reference"] = spot_check(res_ref, rq.records)
                    del bq

        line = {"metric": METRIC, "value": cts_step * args.steps / (ms_dev * 1e-3), "unit": "candidate-timesteps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(world),
                "e2e": {"value": cts_step * args.steps / (ms_e2e * 1e-3), "unit": "candidate-timesteps/s",
                        "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": bp.h2d_bytes * world,
                        "d2h_bytes_per_step": bp.d2h_bytes * world},
                "gpu_launches": bp.launches_per_plan * args.steps,
                "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "parity_spot_check": spot, "clocks": clk,
                "candidate_timesteps_per_step": cts_step,
                "replan_latency": None,
                "literal_obstacle_fields": None, "fp32_fast_path": None,
                "predicted_obstacles": None, "all_feasibility_checks": None, "default_lattice_batch": None, "closed_loop": None,
                "survivors": {"scenarios_with_feasible": int((allrec["status"] == 0).sum()), "scenarios": int(len(allrec))},
                "per_rank": per_rank}
        if world > 1:
            line["config"]["parallelism"] += ("; resident arm: the records of all timed steps all-gathered once at the end of the sweep (inside the "
                                              "timed region)" if not per_step else "; resident arm: all-gather after every step on a side stream")
            line["config"]["parallelism"] += "; e2e arm: all-gather after every step"
        if sizes is not None:
            line["config"]["shard_sizes"] = sizes
            line["config"]["parallelism"] += "; shards sized by each GPU's measured speed (calibration on equal shards, per_rank)"
        rec_f64 = np.array(allrec[:hi - lo], copy=True) if world == 1 else None
        if world == 1 and not args.kernels_only:
            del bp
            torch.cuda.empty_cache()
            # single-replan latency through the drop-in API: the default planner (one kernel per replan with its inputs as the
            # kernel argument, results written to mapped pinned memory, transform / collision stage riding along once the driver
            # has shown its path and a repeating obstacle set), the same without the speculation (what a scene of MOVING obstacles
            # gets), and with every round-2 latency feature off (separate launches in a graph, copy nodes: round 1's path)
            # untimed pass first: single replans leave the GPU idle most of the time, and the first ~10 ms of them after the big legs
            # run 20 % slower than the steady state (scripts/latency_order.py: 38 us for the first leg of a process, 30-32 us after)
            replan_latency(1500)
            lat = replan_latency()
            lat["no_speculation"] = replan_latency(_speculate=False)
            lat["separate_launches_copy_nodes_no_speculation"] = replan_latency(_zero_copy=False, _one_launch=False, _speculate=False)
            line["replan_latency"] = lat
            line["literal_obstacle_fields"] = literal_field_legs(dev, lo, hi, max(3, args.steps // 3))
            line["fp32_fast_path"] = fp32_fast_path(wl, max(3, args.steps // 3), dev, peak, rec_f64)
            line["predicted_obstacles"] = predicted_mode(wl, max(3, args.steps // 3), dev)
            line["all_feasibility_checks"] = all_checks_mode(wl, max(3, args.steps // 3), dev)
            line["default_lattice_batch"] = default_lattice_batch(dev)
            line["closed_loop"] = closed_loop(dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather-per-step", action="store_true", help="N > 1, resident arm: all-gather the records after EVERY step (on the side "
                    "stream) instead of once for the whole sweep of timed steps (BatchPlanner.sweep_begin / sweep_gather, inside the timed "
                    "region): the same on 2 GPUs, 1-2.5 %% slower on 4 and 8 (profiles/r2_history.md section 8).  The e2e arm always "
                    "gathers per step.")
    ap.add_argument("--balance-shards", action="store_true", help="N > 1: size the scenario shards by each GPU's measured speed (short "
                    "calibration on equal shards first) instead of keeping them equal; off by default - see profiles/r2_history.md section 8")
    ap.add_argument("--kernels-only", action="store_true", help="headline step + per-kernel timings only (A/B runs of kernel builds)")
    ap.add_argument("--scenarios", type=int, default=512, help="scenarios per GPU and step (512 = the benchmark; "
                    "smaller only for profiler captures)")
    args = ap.parse_args()
    globals()["SCEN_PER_GPU"] = args.scenarios
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
