"""Batched planning: many independent planner instances ("scenarios") per GPU, scenario-sharded across GPUs.

This is the public API of the batched moving-obstacle workload (BASELINE.json config 5).  One
`BatchPlanner.plan(...)` call is, for every scenario, exactly what one iteration of the reference's
moving-obstacle driver does (lib/tests/test_moving_obstacle_planner.py:86-98):

    planner.replanner(t)            -> K1 + K2 (+ unmasked K5)
    frenet_to_glob(...)             -> K3
    sensor.step_obstacle(t)         -> host time law + device Frenet->Cartesian of the obstacle table
    check_paths(...)                -> K4
    min(paths, key=ctot)            -> K5 masked

Scenarios are independent, so multi-GPU execution is plain scenario-index sharding with a single
gather of the fixed-size result records at the end (`gather_results`); there is no other exchange.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _native as nat
from . import lattice as lat
from .engine import FrenetEngine, LatticeGeometry, make_params


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of scenario indices owned by `rank` (blocks differ by at most one scenario)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_shard_sizes(n_total: int, step_ms, max_dev: float = 0.1) -> list[int]:
    """Shard sizes proportional to each rank's measured speed (1 / its own step time on an equal shard): the GPUs of one box differ
    by several percent on an issue-sensitive kernel, and with equal shards the whole job runs at the slowest one's pace.  Sizes sum
    to n_total and stay within +-max_dev of the equal share; identical inputs give identical sizes on every rank."""
    t = np.asarray(step_ms, dtype=np.float64)
    world = len(t)
    share = n_total / world
    want = np.clip(n_total * (1.0 / t) / np.sum(1.0 / t), share * (1 - max_dev), share * (1 + max_dev))
    sizes = np.floor(want).astype(np.int64)
    order = np.argsort(-(want - sizes), kind="stable")          # largest remainders first
    for i in range(int(n_total - sizes.sum())):
        sizes[order[i % world]] += 1
    while sizes.sum() > n_total:                                 # (clipping at the upper bound can overshoot by rounding)
        sizes[int(np.argmax(sizes))] -= 1
    return [int(v) for v in sizes]


def shard_range_from_sizes(sizes, rank: int) -> tuple[int, int]:
    lo = int(sum(sizes[:rank]))
    return lo, lo + int(sizes[rank])


def gather_results(local: np.ndarray, n_total: int, group=None, device=None, sizes=None) -> np.ndarray:
    """All-gather the per-scenario FrenetResult records of every rank (one collective, NCCL on GPUs, gloo on CPU).

    `local` is this rank's structured record array for its `shard_range` (or for its entry of `sizes`, the shard sizes of all ranks);
    returns all `n_total` records in scenario order on every rank.  Without an initialised process group this is the identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if sizes is not None:
        if len(sizes) != world or sum(sizes) != n_total:
            raise ValueError("sizes must hold one shard size per rank and sum to n_total")
        cap = max(sizes)
        buf = np.zeros(cap * nat.RESULT_BYTES, dtype=np.uint8)
        buf[:local.size * nat.RESULT_BYTES] = local.view(np.uint8).reshape(-1)
        t = torch.from_numpy(buf)
        if device is not None:
            t = t.to(device)
        out = torch.empty(world * cap * nat.RESULT_BYTES, dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
        flat = out.cpu().numpy().reshape(world, cap * nat.RESULT_BYTES)
        return np.concatenate([flat[r, :sizes[r] * nat.RESULT_BYTES].view(np.dtype(nat.RESULT_DTYPE)) for r in range(world)])
    cap = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    buf = np.zeros(cap * nat.RESULT_BYTES, dtype=np.uint8)
    buf[:local.size * nat.RESULT_BYTES] = local.view(np.uint8).reshape(-1)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * cap * nat.RESULT_BYTES, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    flat = out.cpu().numpy().reshape(world, cap * nat.RESULT_BYTES)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(flat[r, :(hi - lo) * nat.RESULT_BYTES].view(np.dtype(nat.RESULT_DTYPE)))
    return np.concatenate(parts)


@dataclass
class BatchResult:
    records: np.ndarray        # FrenetResult per scenario (structured array)
    winner: np.ndarray         # [B][12][n_pad_max] masked winner: t, d, dd, ddd, dddd, s, ds, dds, ddds, x, y, (cd, cv, ct, ctot)
    winner_steps: np.ndarray   # [B]
    cand_timesteps: int        # candidate-timesteps evaluated (sum over valid candidates of their step counts)

    def next_state(self, b: int, time: float, period: float):
        """(s, ds, dds), (d, dd, ddd) of the winner at `time` - optimal_at_time, trajectory_planner.py:183-195."""
        n = int(self.winner_steps[b])
        if n == 0:
            raise ValueError(f"scenario {b}: no feasible candidate")
        w = self.winner[b]
        t = w[0, :n]
        k = 0 if time <= t[0] else (n - 1 if time >= t[-1] else round((time - t[0]) / period))
        return (w[5, k], w[6, k], w[7, k]), (w[1, k], w[2, k], w[3, k])


class BatchPlanner:
    """`n_scen` planner instances with common parameters and lattice geometry on one GPU.

    params:     any object with the reference's planner attributes (e.g. `TrajectoryPlannerParams(...)`).
    variant:    "v1" | "dt" | "dt_obstacles" | "optimal" (which reference class's rules apply).
    trajectory: reference path (CircleTrajectory2D / SplineTrajectory2D), shared by all scenarios.
    """

    def __init__(self, params, trajectory, n_scen: int, n_obs: int = 0, variant: str = "v1", device=None,
                 robot_radius: float = 0.7, predict_steps: int = 0, lat_f32: bool = False, one_launch: bool = False):
        self.p = params
        self.variant = variant
        self.pipeline_slots = 4          # output blocks of launch_pipelined (set before its first call)
        self.n_scen = int(n_scen)
        self.n_obs = int(n_obs)
        self.engine = FrenetEngine(self.n_scen, device)
        self.period = lat.sample_period(params, variant)
        # lat_f32: fp32 FAST PATH (north star: 1e-3) - the sampled lateral / Cartesian state is STORED as floats (half the HBM bytes);
        # arithmetic, costs, collision masks, argmin and the gathered winner stay fp64.  Dense (long-horizon) lattices only.
        self.lat_f32 = bool(lat_f32)
        self.geom = LatticeGeometry.from_intervals(lat.di_interval(params, variant), lat.t_interval(params, variant), self.period,
                                                   8 if self.lat_f32 else 4)
        # predict_steps > 0 (EXTENSION): obstacles are propagated over the horizon on the device and step k of every
        # candidate is tested against the obstacles' predicted positions at t_initial + k * period; 0 = the reference's
        # snapshot semantics (every step against the positions at t_initial).
        self.predict_steps = int(predict_steps)
        self.engine.configure(self.geom, lat.max_long_axis_len(params), self.n_obs, self.predict_steps, self.lat_f32)
        self.path = self.engine.path_handle(trajectory)
        self.radius_sq = robot_radius ** 2
        self.params = make_params(params, lat.VARIANTS[variant]["rule"], self.period, False, self.radius_sq)
        self.stages = nat.STAGE_K1 | nat.STAGE_K2 | nat.STAGE_K3 | nat.STAGE_K5 | nat.STAGE_GATHER
        self.use_mask = 0
        # one_launch: K1 and K5 + gather folded into the evaluate kernel (FRENET_STAGE_ONE_LAUNCH: every CTA solves its own rows'
        # boundary-value systems and reduces its own candidates, the scenario's last CTA finishes) - one kernel per step instead of three
        self.one_launch = bool(one_launch)
        if self.one_launch:
            if self.predict_steps or self.lat_f32:
                raise ValueError("one_launch is not available with predicted obstacles or the fp32 fast path")
            self.stages |= nat.STAGE_ONE_LAUNCH
        if self.n_obs > 0:
            self.stages |= nat.STAGE_K4
            self.use_mask |= nat.MASK_COLLISION
        if np.isfinite(self.params.v_max) or np.isfinite(self.params.a_max):
            self.use_mask |= nat.MASK_KINEMATIC
        if np.isfinite(self.params.kappa_max):
            if self.one_launch:
                raise ValueError("one_launch is not available with the curvature limit")
            self.stages |= nat.STAGE_CURVATURE
            self.use_mask |= nat.MASK_CURVATURE
        self._n_v_sum = 0

    # -- host -> pinned staging ----------------------------------------------------------------
    def stage_lanes(self, lane_fits, projection, road_right=None, road_left=None, ds: float = 1 / 30):
        """Batched Duckietown lane runs (SURVEY section 8f row 2 for many agents): every scenario has its own lane-centre parabola
        `lane_fits[b] = (a, b, c)`, its own projection (the transform evaluates the path at projection + (s - s0[0])) and optionally
        its own lane boundaries `road_right / road_left [B][3]` (rows of NaN = that boundary was not detected).  Call after
        `stage_inputs` (it needs the staged s0); the planner's trajectory must be a ParabolaTrajectory2D (it fixes the path kind)."""
        e, B = self.engine, self.n_scen
        e.set_relative_s(True)
        e.set_per_scenario_lanes(True)
        e.lane_params[:, :3] = np.asarray(lane_fits, dtype=np.float64).reshape(B, 3)
        e.lane_params[:, 3] = ds
        e.path_origin[:, 0] = projection
        e.path_origin[:, 1] = e.state[:, 3]
        e.road_params[:] = 0.0
        any_road = False
        for col, fits in ((0, road_right), (4, road_left)):
            if fits is None:
                continue
            f = np.asarray(fits, dtype=np.float64).reshape(B, 3)
            given = ~np.isnan(f).any(axis=1)
            a, b, c = (np.where(given, f[:, i], 1.0) for i in range(3))
            e.road_params[:, col] = -b / (2 * a)                       # get_vertex_and_focus_distance (test_mapper_planner_obstacles.py:105-109)
            e.road_params[:, col + 1] = -(b ** 2 - 4 * a * c) / (4 * a)
            e.road_params[:, col + 2] = a
            e.road_params[:, col + 3] = given
            any_road = any_road or bool(given.any())
        if any_road:
            self.stages |= nat.STAGE_OFFROAD
            self.use_mask |= nat.MASK_ROAD
        else:
            self.stages &= ~nat.STAGE_OFFROAD
            self.use_mask &= ~nat.MASK_ROAD

    def stage_inputs(self, p0, s0, t0, dd=None, desired_speed=None, obs_s=None, obs_d=None, obs_xy=None, obs_active=None,
                     obs_speed=None, robot_pose=None, sensor=None, targets=None):
        """Write one batch of host inputs into the pinned staging block (no device work yet).

        robot_pose [B][3] + sensor=(range, aperture): the obstacle mask is computed on the device by the cone gate
        (snapshot obstacles only) instead of being passed in `obs_active`.
        targets [B][n_t][3]: follow mode - quintic longitudinal to the (st, dst, ddst) triples, cost ct
        (trajectory_planner.py:131-144); the V axis then holds the target offsets of `si_interval`."""
        e, B = self.engine, self.n_scen
        e.wait_staging_free()      # the previous call's H2D copy may still be reading the pinned block
        p0 = np.asarray(p0, dtype=np.float64).reshape(B, 3)
        s0 = np.asarray(s0, dtype=np.float64).reshape(B, 3)
        e.state[:, :3], e.state[:, 3:] = p0, s0
        e.scal[:, 0] = 0.0 if dd is None else dd
        e.scal[:, 1] = self.p.desired_speed if desired_speed is None else desired_speed
        e.scal[:, 2] = t0
        # per-scenario target-speed axis: the reference's rounding rule (trajectory_planner.py:120), vectorised.
        # np.round is round-half-to-even like Python's round() on floats.
        q, n = self.p.d_d_s, self.p.num_sample
        follow = targets is not None
        self.params.follow = 1 if follow else 0
        if follow:
            e.target[:] = np.asarray(targets, dtype=np.float64).reshape(B, self.geom.n_t, 3)
        centre = np.zeros(B) if follow else s0[:, 1]            # follow: si_interval is centred on 0 (trajectory_planner.py:119)
        lo = np.round(centre - q * n)
        hi = np.round(centre + q * n + q)
        nv = np.ceil((hi - lo) / q).astype(np.int32)
        if int(nv.max()) > e.n_v_cap:
            raise ValueError("longitudinal axis longer than the configured capacity")
        delta = (lo + q) - lo                                             # np.arange's own step: (start + step) - start
        e.axis_v[:] = lo[:, None] + np.arange(e.n_v_cap)[None, :] * delta[:, None]     # == np.arange(lo, hi, q) element-wise
        e.n_v[:] = nv
        self._n_v_sum = int(nv.sum())
        self._obs_frenet = obs_s is not None
        if self.n_obs:
            if obs_s is not None:
                e.obs_sd[0], e.obs_sd[1] = obs_s, obs_d
                e.obs_speed[:] = 0.0 if obs_speed is None else obs_speed
            elif obs_xy is not None:
                e.obs_xy[:] = obs_xy
            e.obs_active[:] = 1 if obs_active is None else obs_active
        self._gate = None
        if robot_pose is not None and sensor is not None:
            e.robot_pose[:] = robot_pose
            self._gate = (float(sensor[0]), float(sensor[1]))

    def launch(self):
        """H2D of the staged inputs, the pipeline, D2H of records + winners; asynchronous on the current stream."""
        e = self.engine
        e.upload()
        self._place()
        e.launch(self.params, self.path, self.stages, self.use_mask, nat.SEL_MASKED)
        e.download(sync=False)

    def _place(self):
        e = self.engine
        if self.n_obs and self._obs_frenet:
            if self.predict_steps:
                e.predict_obstacles(self.path, self.period)
            else:
                e.place_obstacles(self.path)
        if self.n_obs and getattr(self, "_gate", None) and not self.predict_steps:
            e.sensor_gate(*self._gate)

    def launch_resident(self):
        """The pipeline alone on inputs already in HBM (no host traffic)."""
        e = self.engine
        self._place()
        e.launch(self.params, self.path, self.stages, self.use_mask, nat.SEL_MASKED)

    def launch_pipelined(self, n_total: int | None = None, group=None, gather: bool = True, sizes=None):
        """One resident step whose results leave the device on a SIDE stream: the D2H copy of records + winners and - with a
        process group - the one collective of the path (all-gather of the 64-byte records over NCCL) run under the next step's
        kernels instead of in front of them.  `pipeline_slots` output blocks rotate (default 4), so a rank may run that many steps
        minus one ahead of the slowest one before it has to wait for a block - enough to absorb the step-to-step jitter between GPUs;
        `pipeline_wait()` returns the records of the last step enqueued.  gather=False leaves the collective out (this rank's
        records only) - bench.py uses it to separate the collective's cost from the ranks' own spread."""
        import torch.distributed as dist
        e = self.engine
        if not hasattr(self, "_pipe"):
            n = self.pipeline_slots
            e.enable_out_slots(n)
            self._pipe = dict(side=torch.cuda.Stream(device=e.device), done=[None] * n, gbuf=[None] * n, ghost=[None] * n, pad=[None] * n, last=n - 1, sizes=None,
                              slot_mode=[(False, None)] * n)
        P = self._pipe
        slot = (P["last"] + 1) % len(P["done"])
        compute = torch.cuda.current_stream(e.device)
        if P["done"][slot] is not None:
            compute.wait_event(P["done"][slot])       # the side-stream work of the step that used this block two steps ago
        e.select_out_slot(slot)
        self.launch_resident()
        ready = torch.cuda.Event()
        ready.record(compute)
        world = dist.get_world_size(group) if (gather and dist.is_available() and dist.is_initialized()) else 1
        side = P["side"]
        side.wait_event(ready)
        with torch.cuda.stream(side):
            # the collective FIRST (it needs only the 64-byte records): the ranks' handshake must not queue behind this rank's own
            # multi-megabyte download of winner trajectories
            if world > 1:
                if sizes is None and (n_total is None or n_total != world * self.n_scen):
                    raise ValueError("launch_pipelined needs equal shards (n_total = world size x scenarios of this planner) or `sizes`")
                if sizes is not None and (len(sizes) != world or sizes[dist.get_rank(group)] != self.n_scen):
                    raise ValueError("sizes must hold every rank's shard size, this planner's at its own rank")
                o = e._out_offs["result"]
                local = e._out_dev[o:o + self.n_scen * nat.RESULT_BYTES]
                cap = (max(sizes) if sizes is not None else self.n_scen) * nat.RESULT_BYTES
                if P["gbuf"][slot] is None or P["gbuf"][slot].numel() != world * cap:
                    P["gbuf"][slot] = torch.empty(world * cap, dtype=torch.uint8, device=e.device)
                    P["ghost"][slot] = torch.empty(world * cap, dtype=torch.uint8, pin_memory=True)
                    P["pad"][slot] = torch.zeros(cap, dtype=torch.uint8, device=e.device) if cap != local.numel() else None
                if P["pad"][slot] is not None:         # unequal shards: every rank contributes the largest shard's size
                    P["pad"][slot][:local.numel()].copy_(local, non_blocking=True)
                    local = P["pad"][slot]
                dist.all_gather_into_tensor(P["gbuf"][slot], local, group=group)
                P["ghost"][slot].copy_(P["gbuf"][slot], non_blocking=True)
                P["sizes"] = list(sizes) if sizes is not None else None
            sw = getattr(self, "_sweep", None)
            if sw is not None and sw["at"] < sw["steps"]:      # sweep: this step's records stay on the device until sweep_gather
                o = e._out_offs["result"]
                sw["dev"][sw["at"]].copy_(e._out_dev[o:o + self.n_scen * nat.RESULT_BYTES], non_blocking=True)
                sw["at"] += 1
            e._out_host.copy_(e._out_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(side)
        P["done"][slot] = done
        P["last"] = slot
        P["gathered"] = world > 1
        P["slot_mode"][slot] = (world > 1, list(sizes) if (sizes is not None and world > 1) else None)     # what THIS block's step did

    def launch_streamed(self, n_total: int | None = None, group=None, sizes=None) -> int:
        """`launch()` whose results leave on the side stream: H2D of the staged inputs and the pipeline on the current stream, the
        all-gather (with a process group) and the D2H of records + winners behind an event on the pipeline's side stream, into the
        next of the rotating output blocks.  Returns the block's slot for `collect(slot)`.  A host loop that stages step i + 1 and
        collects step i - 1 while step i runs keeps PCIe, host and kernels busy at the same time:

            slot = bp.launch_streamed(); bp.stage_inputs(next...); res = bp.collect(previous_slot); previous_slot = slot"""
        self.engine.upload()
        self.launch_pipelined(n_total, group, True, sizes)
        return self._pipe["last"]

    def collect(self, slot: int) -> BatchResult:
        """Wait for the side-stream work of the step that wrote output block `slot` and hand out its results (views of that block's
        pinned host copy, valid until the block comes round again: `pipeline_slots` launches later).  With a process group `records`
        are all ranks' (scenario order), winners this rank's."""
        P, e = self._pipe, self.engine
        P["done"][slot].synchronize()
        rec, win, steps = e.out_views(slot)
        gathered, sz = P["slot_mode"][slot]
        if gathered:
            flat = P["ghost"][slot].numpy()
            if sz is None:
                rec = flat.view(np.dtype(nat.RESULT_DTYPE))
            else:
                rows = flat.reshape(len(sz), -1)
                rec = np.concatenate([rows[r, :sz[r] * nat.RESULT_BYTES].view(np.dtype(nat.RESULT_DTYPE)) for r in range(len(sz))])
        return BatchResult(rec, win, steps, self.cand_timesteps())

    def sweep_begin(self, n_steps: int):
        """A SWEEP of `n_steps` pipelined steps whose records are gathered ONCE at its end: every following
        `launch_pipelined(gather=False)` also keeps its step's records in a device block [n_steps][n_scen]; `sweep_gather` then
        runs the path's one collective over all of them.  Between the GPUs nothing is exchanged (and nobody waits for anybody)
        until then - a per-step collective paces every rank at the slowest one's step and keeps an NCCL kernel spinning next to
        the pipeline kernel."""
        e = self.engine
        n = int(n_steps)
        sw = getattr(self, "_sweep", None)
        if sw is None or sw["steps"] != n:
            sw = dict(steps=n, dev=torch.empty((n, self.n_scen * nat.RESULT_BYTES), dtype=torch.uint8, device=e.device), gbuf=None, ghost=None)
            self._sweep = sw
        sw["at"] = 0

    def sweep_gather(self, n_total: int, group=None, sync: bool = True):
        """The one collective of a sweep (equal shards): all ranks' records of all its steps, [n_steps][n_total] in scenario order.
        Enqueued on the pipeline's side stream behind the last step; sync=False returns None (call `sweep_records()` later)."""
        import torch.distributed as dist
        sw, P, e = self._sweep, self._pipe, self.engine
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if world > 1 and n_total != world * self.n_scen:
            raise ValueError("sweep_gather needs equal shards: n_total = world size x scenarios of this planner")
        with torch.cuda.stream(P["side"]):
            flat = sw["dev"].reshape(-1)
            if world > 1:
                if sw["gbuf"] is None:
                    sw["gbuf"] = torch.empty(world * flat.numel(), dtype=torch.uint8, device=e.device)
                dist.all_gather_into_tensor(sw["gbuf"], flat, group=group)
                flat = sw["gbuf"]
            if sw["ghost"] is None or sw["ghost"].numel() != flat.numel():
                sw["ghost"] = torch.empty(flat.numel(), dtype=torch.uint8, pin_memory=True)
            sw["ghost"].copy_(flat, non_blocking=True)
            sw["done"] = torch.cuda.Event()
            sw["done"].record(P["side"])
        sw["world"] = world
        P["done"][P["last"]] = sw["done"]          # pipeline_join / pipeline_wait now also cover the sweep's collective
        return self.sweep_records() if sync else None

    def sweep_records(self):
        sw = self._sweep
        sw["done"].synchronize()
        rec = sw["ghost"].numpy().view(np.dtype(nat.RESULT_DTYPE)).reshape(sw["world"], sw["steps"], self.n_scen)
        return np.ascontiguousarray(rec.transpose(1, 0, 2)).reshape(sw["steps"], sw["world"] * self.n_scen)

    def pipeline_join(self):
        """Make the current stream wait for all side-stream work enqueued so far (for timing the whole pipeline with events)."""
        P = getattr(self, "_pipe", None)
        if P:
            for ev in P["done"]:
                if ev is not None:
                    torch.cuda.current_stream(self.engine.device).wait_event(ev)

    def pipeline_wait(self):
        """Records of the last pipelined step: all ranks' (scenario order) with a process group, else this planner's."""
        P = self._pipe
        P["done"][P["last"]].synchronize()
        if P.get("gathered"):
            flat = P["ghost"][P["last"]].numpy()
            if P.get("sizes") is None:
                return flat.view(np.dtype(nat.RESULT_DTYPE))
            sz = P["sizes"]
            rows = flat.reshape(len(sz), -1)
            return np.concatenate([rows[r, :sz[r] * nat.RESULT_BYTES].view(np.dtype(nat.RESULT_DTYPE)) for r in range(len(sz))])
        return self.engine.result

    def enable_follow_fallback(self, obs_time_offset, obs_offset, obs_speed, d_target: float = 2.0):
        """Batched form of the moving-obstacle driver's all-blocked branch (lib/tests/test_moving_obstacle_planner.py:92-97), on the
        device: a scenario whose candidates all collide re-plans in follow mode behind `leading_obstacles[0]` - the first blocker of
        its first candidate - and, like the reference planner whose `s_target` stays set, keeps planning in follow mode afterwards.
        obs_time_offset / obs_offset / obs_speed [B][O]: the obstacles' ObstacleTrajectory parameters (time law
        `time_offset + (t * speed) % 50`, lateral offset; lib/sensor/dynamic_obstacle.py:36-56) that `compute_s / ds / dds` need;
        d_target: the planner's `d_target` (2.0 in `initialize`, trajectory_planner.py:91)."""
        e = self.engine
        e.enable_follow()
        e.follow_dev.zero_()
        e.lead_dev.fill_(-1)
        e.obs_sd[0], e.obs_sd[1] = obs_time_offset, obs_offset
        e.obs_speed[:] = obs_speed
        e.upload_region("obs_sd", "obs_speed")
        self._follow = dict(d_target=float(d_target))

    def pipeline_with_follow(self, estimate):
        """The pipeline on the state resident in HBM, then - with `enable_follow_fallback` - the follow re-plan of the scenarios
        that came out without a survivor, all on the device.  estimate: device tensor [B], arc length of every robot's projection
        (what `transformer.estimatePosition` returned this step): it anchors the frame the follow targets are measured in."""
        e = self.engine
        fb = getattr(self, "_follow", None)
        if fb is not None:
            e.follow_targets(self.path, estimate, self.p.delta_t, fb["d_target"])          # scenarios that already follow
        e.launch(self.params, self.path, self.stages, self.use_mask, nat.SEL_MASKED)
        if fb is not None and (self.stages & nat.STAGE_K4):
            e.follow_select(self.p.d_d_s, self.p.num_sample)
            e.follow_targets(self.path, estimate, self.p.delta_t, fb["d_target"], only=e.redo_dev)
            e.set_scenario_mask(e.redo_dev)
            # the reference takes min(paths, key=ctot) over the UNFILTERED follow lattice (it still runs frenet_to_glob on it)
            e.launch(self.params, self.path, self.stages & ~nat.STAGE_K4, 0, nat.SEL_MASKED)
            e.set_scenario_mask(None)

    def replanner(self, time, robot_pose=None, sensor=None, obs_xy=None, sync: bool = True, estimate=None):
        """One `planner.replanner(time)` for every scenario WITHOUT a host round trip for the planner state
        (trajectory_planner.py:206-212, 223-234): the new initial state is the previous (masked) winner sampled at `time`,
        taken on the device; then the pipeline runs again.  Call `plan(...)` once first (it plays the role of
        `initialize`).  robot_pose + sensor=(range, aperture) re-evaluate the cone gate; obs_xy moves the obstacles.
        estimate [B] (with `enable_follow_fallback`): arc length of the robots' projections, see `pipeline_with_follow`."""
        e = self.engine
        regions = []
        if robot_pose is not None:
            e.robot_pose[:] = robot_pose
            regions.append("robot_pose")
        if obs_xy is not None:
            e.obs_xy[:] = obs_xy
            regions.append("obs_xy")
        if regions:
            e.upload_region(*regions)
        lookup = getattr(self.p, "sampling_t", None) if self.variant == "dt_obstacles" else self.p.delta_t
        e.advance_state(time, lookup, self.p.d_d_s, self.p.num_sample, follow=bool(self.params.follow))
        if self.n_obs and sensor is not None:
            e.sensor_gate(float(sensor[0]), float(sensor[1]))
        if getattr(self, "_follow", None) is not None:
            if estimate is None:
                raise ValueError("the follow fallback needs the robots' projections (estimate=)")
            if not isinstance(estimate, torch.Tensor):
                estimate = torch.as_tensor(np.ascontiguousarray(estimate, dtype=np.float64)).to(e.device)
            self.pipeline_with_follow(estimate)
        else:
            e.launch(self.params, self.path, self.stages, self.use_mask, nat.SEL_MASKED)
        e.download(sync=sync)
        if sync:
            self._n_v_sum = int(e.result["n_valid"].sum()) // max(self.geom.n_t * self.geom.n_d, 1)
        return e.result

    def gather_records(self, n_total: int, group=None, sync: bool = True, sizes=None):
        """All ranks' result records of the last launch, in scenario order: ONE NCCL all-gather straight from the device
        block K5 wrote (no host staging), then one copy to the host.  Must follow a launch on this stream; shard sizes
        follow `shard_range`.  Without a process group it returns this planner's own records.
        sync=False only enqueues the collective and the copy (equal shards) and returns None: the records are in the array a later
        synchronising call returns, once the stream has been synchronised."""
        import torch.distributed as dist
        e = self.engine
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self._local_records()
        world = dist.get_world_size(group)
        if sizes is not None:       # shards sized by the caller (e.g. by measured GPU speed): padded path, every rank alike
            if not sync:
                raise ValueError("gather_records(sync=False) needs equal shards")
            return gather_results(self._local_records(), n_total, group, self._collective_device(group), sizes=sizes)
        if n_total % world != 0:    # ragged shards (decided from n_total alone, so EVERY rank takes this branch): padded host path
            if not sync:
                raise ValueError("gather_records(sync=False) needs equal shards (n_total divisible by the world size)")
            return gather_results(self._local_records(), n_total, group, self._collective_device(group))
        if n_total // world != self.n_scen:
            raise ValueError(f"this planner holds {self.n_scen} scenarios, its shard of {n_total} over {world} ranks has {n_total // world}")
        o = e._out_offs["result"]
        local = e._out_dev[o:o + self.n_scen * nat.RESULT_BYTES]
        if getattr(self, "_gather_buf", None) is None or self._gather_buf.numel() != world * local.numel():
            self._gather_buf = torch.empty(world * local.numel(), dtype=torch.uint8, device=e.device)
            self._gather_host = torch.empty(world * local.numel(), dtype=torch.uint8, pin_memory=True)
        dist.all_gather_into_tensor(self._gather_buf, local, group=group)
        self._gather_host.copy_(self._gather_buf, non_blocking=True)
        if not sync:
            return None
        torch.cuda.current_stream(e.device).synchronize()
        return self._gather_host.numpy().view(np.dtype(nat.RESULT_DTYPE))

    def _local_records(self):
        """This planner's records of the last launch, once the stream has drained."""
        torch.cuda.current_stream(self.engine.device).synchronize()
        return self.engine.result

    def _collective_device(self, group=None):
        import torch.distributed as dist
        return self.engine.device if dist.get_backend(group) == "nccl" else None

    @property
    def launches_per_plan(self) -> int:
        """Kernel launches of one `launch()` (what bench.py reports as gpu_launches)."""
        st = self.stages
        kernels = [nat.STAGE_K1, nat.STAGE_K2, nat.STAGE_K3, nat.STAGE_CURVATURE, nat.STAGE_K4, nat.STAGE_OFFROAD, nat.STAGE_K5, nat.STAGE_GATHER]
        n = sum(1 for k in kernels if st & k)
        if not (st & nat.STAGE_NO_FUSE):
            if (st & nat.STAGE_K2) and (st & nat.STAGE_K3):     # frenet_plan fuses K2 + K3 (+ K4, + curvature)
                n -= 1 + (1 if st & nat.STAGE_K4 else 0) + (1 if st & nat.STAGE_CURVATURE else 0)
            if (st & nat.STAGE_K5) and (st & nat.STAGE_GATHER):  # argmin + winner gather
                n -= 1
        if st & nat.STAGE_ONE_LAUNCH:
            n = 1                                                # K1, evaluate (+ K3, K4), K5 + gather: one kernel
        if self.predict_steps and (st & nat.STAGE_K4):
            n += 1                                               # per-chunk obstacle bounds
        return n + (1 if (self.n_obs and getattr(self, "_obs_frenet", False)) else 0)

    @property
    def h2d_bytes(self) -> int:
        return self.engine._in_host.numel()

    @property
    def d2h_bytes(self) -> int:
        return self.engine._out_host.numel()

    def cand_timesteps(self) -> int:
        return int(self.geom.n_steps.astype(np.int64).sum()) * self.geom.n_d * self._n_v_sum

    def result(self, copy: bool = False) -> BatchResult:
        """Wait for the last launch and hand out its results.  The arrays are views of the pinned block the device wrote
        into - valid until the next launch of this planner; pass copy=True (or call `.copy()` on them) to keep them."""
        torch.cuda.current_stream(self.engine.device).synchronize()
        e = self.engine
        if copy:
            return BatchResult(e.result.copy(), e.winner.copy(), e.winner_steps.copy(), self.cand_timesteps())
        return BatchResult(e.result, e.winner, e.winner_steps, self.cand_timesteps())

    def plan(self, p0, s0, t0, dd=None, desired_speed=None, obs_s=None, obs_d=None, obs_xy=None, obs_active=None,
             obs_speed=None, robot_pose=None, sensor=None, targets=None) -> BatchResult:
        """Host arrays in, per-scenario results out (synchronous)."""
        self.stage_inputs(p0, s0, t0, dd, desired_speed, obs_s, obs_d, obs_xy, obs_active, obs_speed, robot_pose, sensor, targets)
        self.launch()
        return self.result(copy=True)


def plan_sharded(params, trajectory, scenarios: dict, variant: str = "v1", tile: int = 512, group=None):
    """Plan all scenarios of `scenarios` (dict of arrays with a leading scenario axis: p0, s0, t0, dd and
    obs_s/obs_d or obs_xy) on this process's GPU shard and gather every rank's records.

    Returns (records for ALL scenarios in index order, BatchResult list of this rank's tiles)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n_total = len(scenarios["t0"])
    lo, hi = shard_range(n_total, rank, world)
    n_obs = scenarios["obs_s"].shape[1] if "obs_s" in scenarios else (scenarios["obs_xy"].shape[1] if "obs_xy" in scenarios else 0)
    planner = None
    tiles, recs = [], []
    for a in range(lo, hi, tile):
        b = min(a + tile, hi)
        if planner is None or planner.n_scen != b - a:
            planner = BatchPlanner(params, trajectory, b - a, n_obs, variant)
        kw = {k: scenarios[k][a:b] for k in ("dd", "obs_s", "obs_d", "obs_xy", "obs_active") if k in scenarios}
        res = planner.plan(scenarios["p0"][a:b], scenarios["s0"][a:b], scenarios["t0"][a:b], **kw)
        tiles.append(res)
        recs.append(res.records)
    local = np.concatenate(recs) if recs else np.zeros(0, dtype=np.dtype(nat.RESULT_DTYPE))
    dev = torch.device("cuda", torch.cuda.current_device()) if world > 1 and dist.get_backend(group) == "nccl" else None
    return gather_results(local, n_total, group, dev), tiles
