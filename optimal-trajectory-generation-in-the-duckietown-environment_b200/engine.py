"""Device-side engine: owns the HBM buffers of a batch of planner instances and drives the C ABI.

Everything numeric happens in `libfrenet_b200.so` (csrc/frenet_kernels.cu).  This module only
  * turns lattice intervals into axes and step counts with the same NumPy calls the reference uses
    (`np.arange`, SURVEY.md section 7.3 "bit-faithful lattice"),
  * lays the buffers out in HBM (include/frenet_b200.h documents the layout),
  * stages the small per-call inputs in ONE pinned host block and the small per-call outputs
    (result records + winner trajectories) in another, so a replan costs one H2D copy, one
    `frenet_plan` call and one D2H copy,
  * hands NumPy views of device data back on request.

PyTorch is used for device memory, pinned memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _native as nat
from .trajectory.paths import path_descriptor_of, PATH_KIND_SPLINE


# Every sampled field is padded to a multiple of PAD_STEPS doubles = one 32-byte sector; the kernels write the
# padding samples too, so no sector is ever written partially (measured: a partially written sector per field
# costs a DRAM read-modify-write and caps the store bandwidth at ~4.1 TB/s instead of ~7 TB/s).
PAD_STEPS = 4


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("the Frenet planner hot path runs on a CUDA device only (sm_100a); "
                           "no CUDA device is visible and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class LatticeGeometry:
    """Axes shared by every scenario of a batch: lateral offsets D, horizons T, step counts.

    `np.arange(0, T + period, period)` defines the samples of a horizon
    (lib/planner/trajectory_planner.py:139,148); only its length is needed on the device because
    sample k of that array is exactly `k * period`.
    """

    def __init__(self, D, T, period: float, pad: int = PAD_STEPS):
        self.D = np.ascontiguousarray(D, dtype=np.float64)
        self.T = np.ascontiguousarray(T, dtype=np.float64)
        self.period = float(period)
        self.pad = int(pad)          # 4 doubles = one 32-byte sector; 8 for the fp32 fast path (8 floats)
        self.n_steps = np.array([len(np.arange(0, Tj + period, period)) for Tj in self.T], dtype=np.int32)
        npad = (self.n_steps + self.pad - 1) // self.pad * self.pad
        self.n_pad = npad.astype(np.int32)
        self.t_offset = np.concatenate([[0], np.cumsum(npad)]).astype(np.int32)
        self.n_t, self.n_d = len(self.T), len(self.D)
        self.n_pad_max = int(npad.max()) if self.n_t else 0
        self.sum_pad = int(self.t_offset[-1])
        # t_k ** n exactly as the reference forms them (Python float pow), k = 0 .. n_pad_max - 1
        self.t_pow = np.array([[float(np.float64(k) * np.float64(self.period)) ** n for n in (2, 3, 4, 5)]
                               for k in range(self.n_pad_max)], dtype=np.float64).reshape(self.n_pad_max, 4)
        self.key = (self.D.tobytes(), self.T.tobytes(), self.period, self.pad)

    @staticmethod
    def from_intervals(di_interval, t_interval, period, pad: int = PAD_STEPS):
        return LatticeGeometry(np.arange(*di_interval), np.arange(*t_interval), period, pad)

    def steps_of_candidates(self, n_v: int) -> np.ndarray:
        """Step count of every candidate in list order for an axis of n_v longitudinal targets."""
        return np.tile(np.repeat(self.n_steps, self.n_d), n_v)

    def cand_timesteps(self, n_v) -> int:
        return int(np.sum(self.n_steps.astype(np.int64)) * self.n_d * np.sum(n_v))


@dataclass
class PathHandle:
    desc: nat.FrenetPath
    table: torch.Tensor | None    # keeps the device spline table alive
    kind: int


def make_path_handle(trajectory, device) -> PathHandle:
    kind, n_knots, block, table = path_descriptor_of(trajectory)
    desc = nat.FrenetPath()
    desc.kind = kind
    desc.n_knots = n_knots
    for i in range(16):
        desc.circle[i] = float(block[i])
    tab = None
    if kind == PATH_KIND_SPLINE:
        tab = torch.from_numpy(np.ascontiguousarray(table, dtype=np.float64)).to(device)
        desc.spline_table = tab.data_ptr()
    return PathHandle(desc, tab, kind)


def make_params(p, variant_rule: int, sample_period: float, follow: bool, radius_sq: float = 0.7 ** 2) -> nat.FrenetParams:
    """FrenetParams from any object carrying the reference's planner attributes."""
    P = nat.FrenetParams()
    for k in ("kj", "kt", "kd", "ks", "kdots", "klat", "klong", "low_speed_threshold"):
        setattr(P, k, float(getattr(p, k)))
    P.sample_period = float(sample_period)
    P.s_thresh = float(getattr(p, "S_thresh", 0.0))
    P.v_max = float(getattr(p, "v_max", math.inf))
    P.a_max = float(getattr(p, "a_max", math.inf))
    P.kappa_max = float(getattr(p, "kappa_max", math.inf))
    P.radius_sq = float(radius_sq)
    P.lowspeed_rule = int(variant_rule)
    P.follow = 1 if follow else 0
    return P


def _align(n, a=256):
    return (n + a - 1) // a * a


def _on_device(fn):
    """Run an engine method with the engine's GPU as the current CUDA device: the kernels are `<<<>>>` launches on the CURRENT
    device, and attribute setting, pinned allocation and graph capture follow it too.  One integer compare when it already is."""
    import functools

    get_device = torch._C._cuda_getDevice      # the bare runtime call: torch.cuda.current_device() adds ~3 us of lazy-init checks per call

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        if get_device() == self._dev_index:
            return fn(self, *a, **k)
        with torch.cuda.device(self._dev_index):
            return fn(self, *a, **k)
    return wrapper


class FrenetEngine:
    """HBM buffers + launch logic for `n_scen` independent planner instances sharing one lattice geometry."""

    def __init__(self, n_scen: int = 1, device=None, zero_copy: bool = False, masked_winner: bool = False):
        """zero_copy: the packed per-call inputs and outputs are NOT mirrored in HBM - the kernels read the pinned host block and
        write the result record + winner trajectory straight into pinned host memory over PCIe (under unified addressing a pinned
        allocation has the same address on the device).  For ONE planner these are a few hundred bytes in and a few KB out, and a
        replan is launch-bound.  "out" maps only the outputs (posted writes, nothing waits for them); True maps the inputs too
        (every CTA then pays a PCIe round trip for its first loads - measured in profiles/r2_history.md, latency section).
        Batches keep device blocks (their state stays resident)."""
        if zero_copy not in (False, True, "out"):
            raise ValueError('zero_copy must be False, True or "out"')
        self.zero_copy = zero_copy        # "out": only the result record + winner block are written straight to pinned host memory
        self._zc_in = zero_copy is True
        self._zc_out = bool(zero_copy)
        # masked_winner: a second winner block that the one-launch replan fills with the MASKED pick (FrenetBuffers.winner_masked)
        self.masked_winner = bool(masked_winner)
        self._pending = False         # zero_copy: a launch was enqueued and not waited for yet (it may still read the pinned inputs)
        self.device = require_cuda() if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FrenetEngine needs a CUDA device")
        self.n_scen = int(n_scen)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.geom: LatticeGeometry | None = None
        self.n_v_cap = 0
        self.n_obs = 0
        self.L = nat.FrenetLattice()
        self.B = nat.FrenetBuffers()
        self._cap_key = None
        self._graphs = {}
        self._paths = {}
        self.generation = 0
        self.has_xy = False
        self.has_collision = False
        self.xy_path = None           # the trajectory object the lattice's x, y were computed against (drivers.frenet_to_glob / planner)
        self.spec_pos = None          # obstacle positions the last replan tested speculatively (drivers.collision_filter validates)
        self.relative_s = False       # Duckietown lane runs: the transform evaluates the path at path_origin + (s - s0[0])
        self.per_scenario_lanes = False   # every scenario has its own lane-centre parabola and lane boundaries (lane_params, road_params)

    # -- allocation ---------------------------------------------------------------------------
    @_on_device
    def configure(self, geom: LatticeGeometry, n_v_cap: int, n_obs: int = 0, n_obs_steps: int = 0, lat_f32: bool = False):
        """(Re)allocate for a geometry / capacities; cheap when nothing changed.
        n_obs_steps > 0 additionally allocates a predicted obstacle table [B][n_obs][n_obs_steps][2] (extension).
        lat_f32: fp32 FAST PATH - the candidate blocks (d and its derivatives, x, y) are stored as floats (FrenetLattice.lat_f32);
        the geometry must then be padded to 8 samples."""
        n_obs, n_obs_steps = int(n_obs), int(n_obs_steps)
        lat_f32 = bool(lat_f32)
        if lat_f32 and geom.pad % 8:
            raise ValueError("the fp32 fast path needs a geometry padded to 8 samples (LatticeGeometry(..., pad=8))")
        key = (geom.key, int(n_v_cap), n_obs, n_obs_steps, lat_f32)
        if key == self._cap_key:
            return
        dev, Bn = self.device, self.n_scen
        self._graphs.clear()          # captured graphs point into the buffers that are about to be replaced
        # ... and so do views of the previous lattice: they go stale now, not at the next launch
        self.generation += 1
        self.has_xy = self.has_collision = False
        self.xy_path = self.spec_pos = None
        self.geom, self.n_v_cap, self.n_obs = geom, int(n_v_cap), n_obs
        nT, nD = geom.n_t, geom.n_d
        R = self.n_v_cap * nT
        Cc = R * nD
        self.R, self.C = R, Cc
        row_elems = self.n_v_cap * geom.sum_pad
        cand_elems = row_elems * nD
        self.row_elems, self.cand_elems = row_elems, cand_elems

        # geometry arrays (one small upload per geometry change)
        g = np.zeros(_align(nD * 8) + _align(nT * 8) + _align(nT * 4) + _align((nT + 1) * 4) + _align(geom.n_pad_max * 32),
                     dtype=np.uint8)
        o_d, o_t = 0, _align(nD * 8)
        o_n = o_t + _align(nT * 8)
        o_o = o_n + _align(nT * 4)
        o_p = o_o + _align((nT + 1) * 4)
        g[o_p:o_p + geom.n_pad_max * 32] = geom.t_pow.reshape(-1).view(np.uint8)
        g[o_d:o_d + nD * 8] = geom.D.view(np.uint8)
        g[o_t:o_t + nT * 8] = geom.T.view(np.uint8)
        g[o_n:o_n + nT * 4] = geom.n_steps.view(np.uint8)
        g[o_o:o_o + (nT + 1) * 4] = geom.t_offset.view(np.uint8)
        self._geom_dev = torch.from_numpy(g).to(dev)
        gp = self._geom_dev.data_ptr()

        # packed small inputs: one pinned host block, one device block
        sizes = [("state", Bn * 6 * 8), ("scal", Bn * 3 * 8), ("axis_v", Bn * self.n_v_cap * 8), ("target", Bn * nT * 3 * 8),
                 ("obs_xy", Bn * max(n_obs, 1) * 2 * 8), ("obs_sd", Bn * max(n_obs, 1) * 2 * 8),
                 ("obs_speed", Bn * max(n_obs, 1) * 8), ("robot_pose", Bn * 3 * 8), ("path_origin", Bn * 16), ("lane_params", Bn * 32), ("road_params", Bn * 64),
                 ("n_v", Bn * 4), ("obs_active", Bn * max(n_obs, 1))]
        offs, tot = {}, 0
        for name, sz in sizes:
            offs[name] = tot
            tot += _align(sz)
        self._in_host = torch.empty(tot, dtype=torch.uint8, pin_memory=True)
        self._in_dev = self._in_host if self._zc_in else torch.empty(tot, dtype=torch.uint8, device=dev)
        hv = self._in_host.numpy()
        self._in_host_bytes = hv          # writable byte view of the staging block (struct.pack_into target of the planner's fast path)
        self.state = hv[offs["state"]:offs["state"] + Bn * 48].view(np.float64).reshape(Bn, 6)
        self.scal = hv[offs["scal"]:offs["scal"] + Bn * 24].view(np.float64).reshape(Bn, 3)
        self.axis_v = hv[offs["axis_v"]:offs["axis_v"] + Bn * self.n_v_cap * 8].view(np.float64).reshape(Bn, self.n_v_cap)
        self.target = hv[offs["target"]:offs["target"] + Bn * nT * 24].view(np.float64).reshape(Bn, nT, 3)
        no = max(n_obs, 1)
        self.obs_xy = hv[offs["obs_xy"]:offs["obs_xy"] + Bn * no * 16].view(np.float64).reshape(Bn, no, 2)
        self.obs_sd = hv[offs["obs_sd"]:offs["obs_sd"] + Bn * no * 16].view(np.float64).reshape(2, Bn, no)   # [s|d][B][O]
        self.obs_speed = hv[offs["obs_speed"]:offs["obs_speed"] + Bn * no * 8].view(np.float64).reshape(Bn, no)
        self.robot_pose = hv[offs["robot_pose"]:offs["robot_pose"] + Bn * 24].view(np.float64).reshape(Bn, 3)
        self.path_origin = hv[offs["path_origin"]:offs["path_origin"] + Bn * 16].view(np.float64).reshape(Bn, 2)
        self.lane_params = hv[offs["lane_params"]:offs["lane_params"] + Bn * 32].view(np.float64).reshape(Bn, 4)
        self.road_params = hv[offs["road_params"]:offs["road_params"] + Bn * 64].view(np.float64).reshape(Bn, 8)
        self.n_v = hv[offs["n_v"]:offs["n_v"] + Bn * 4].view(np.int32)
        self.obs_active = hv[offs["obs_active"]:offs["obs_active"] + Bn * no].reshape(Bn, no)
        hv[:] = 0
        self.obs_active[:] = 1
        ip = self._in_dev.data_ptr()
        self._in_offs = offs

        # packed small outputs
        osz = [("winner", Bn * 12 * geom.n_pad_max * 8), ("result", Bn * nat.RESULT_BYTES), ("winner_steps", Bn * 4)]
        if self.masked_winner:
            osz += [("winner_masked", Bn * 12 * geom.n_pad_max * 8), ("winner_masked_steps", Bn * 4)]
        ooffs, otot = {}, 0
        for name, sz in osz:
            ooffs[name] = otot
            otot += _align(sz)
        self._out_offs = ooffs
        self._out_bytes = otot
        self._out_blocks = []          # (pinned host block, device block); a second one is added by enable_out_slots(2)
        self._add_out_block()
        self._bind_out_block(0)
        op = self._out_dev.data_ptr()

        # large state
        f64 = dict(dtype=torch.float64, device=dev)
        self.coef_lon = torch.empty((Bn, R, 6), **f64)
        self.coef_lat = torch.empty((Bn, Cc, 6), **f64)
        self.row_S = torch.empty((Bn, R), **f64)
        self.row_flags = torch.empty((Bn, R), dtype=torch.int32, device=dev)
        self.lon = torch.empty((Bn, row_elems * 4), **f64)     # per row:       [4][npad]  s, ds, dds, ddds
        self.lat_f32 = lat_f32
        self.lat = torch.empty((Bn, cand_elems * 6), dtype=torch.float32 if lat_f32 else torch.float64, device=dev)   # per candidate: [6][npad]  d, dd, ddd, dddd, x, y
        self.cost = torch.empty((4, Bn, Cc), **f64)
        self.cand_flags = torch.empty((Bn, Cc), dtype=torch.uint8, device=dev)
        self.curv_ok = torch.empty((Bn, Cc), dtype=torch.uint8, device=dev)
        self.coll_free = torch.empty((Bn, Cc), dtype=torch.uint8, device=dev)
        self.first_blocker = torch.empty((Bn, Cc), dtype=torch.int32, device=dev)
        self.road_ok = torch.empty((Bn, Cc), dtype=torch.uint8, device=dev)
        # FRENET_STAGE_ONE_LAUNCH scratch: completion tickets per scenario, a partial argmin record per CTA
        self.tickets = torch.zeros(Bn, dtype=torch.int32, device=dev)
        # (a row is split over several CTAs only while all of them fit the GPU at once - a handful of scenarios; beyond that one
        #  record per row is the most a scenario's CTAs can need)
        parts_cap = nat.MAX_ROW_PARTS if Bn * R <= 4 * 148 else 1
        self.partials = torch.empty(Bn * R * parts_cap * nat.PARTIAL_BYTES, dtype=torch.uint8, device=dev)
        self.n_obs_steps = n_obs_steps
        self.obs_xy_t = self.obs_bounds = None
        if n_obs_steps > 0 and n_obs > 0:
            self.obs_xy_t = torch.empty((Bn, n_obs, n_obs_steps, 2), **f64)
            self.obs_bounds = torch.empty((Bn, (geom.n_pad_max + 31) // 32, n_obs, 4), dtype=torch.float32, device=dev)

        L = self.L
        L.n_scen, L.n_v_cap, L.n_t, L.n_d, L.n_pad_max = Bn, self.n_v_cap, nT, nD, geom.n_pad_max
        L.lat_f32 = 1 if lat_f32 else 0
        L.row_elems, L.cand_elems = row_elems, cand_elems
        L.axis_d, L.axis_t, L.n_steps, L.t_offset, L.t_pow = gp + o_d, gp + o_t, gp + o_n, gp + o_o, gp + o_p
        L.axis_v, L.n_v = ip + offs["axis_v"], ip + offs["n_v"]
        Bf = self.B
        Bf.state, Bf.scal, Bf.target = ip + offs["state"], ip + offs["scal"], ip + offs["target"]
        Bf.obs_xy, Bf.obs_active, Bf.n_obs = ip + offs["obs_xy"], ip + offs["obs_active"], n_obs
        Bf.coef_lon, Bf.coef_lat = self.coef_lon.data_ptr(), self.coef_lat.data_ptr()
        Bf.row_S, Bf.row_flags = self.row_S.data_ptr(), self.row_flags.data_ptr()
        Bf.lon, Bf.lat, Bf.cost, Bf.cand_flags = self.lon.data_ptr(), self.lat.data_ptr(), self.cost.data_ptr(), self.cand_flags.data_ptr()
        Bf.curv_ok = self.curv_ok.data_ptr()
        Bf.coll_free, Bf.first_blocker = self.coll_free.data_ptr(), self.first_blocker.data_ptr()
        Bf.n_obs_steps = n_obs_steps if self.obs_xy_t is not None else 0
        Bf.obs_xy_t = self.obs_xy_t.data_ptr() if self.obs_xy_t is not None else None
        Bf.obs_bounds = self.obs_bounds.data_ptr() if self.obs_bounds is not None else None
        Bf.result = op + ooffs["result"]
        Bf.winner, Bf.winner_steps = op + ooffs["winner"], op + ooffs["winner_steps"]
        Bf.road_ok = self.road_ok.data_ptr()
        Bf.tickets = self.tickets.data_ptr()
        Bf.partials = self.partials.data_ptr() if self.partials is not None else None
        Bf.path_origin = ip + offs["path_origin"] if self.relative_s else None
        Bf.lane_params = ip + offs["lane_params"] if self.per_scenario_lanes else None
        Bf.road_params = ip + offs["road_params"] if self.per_scenario_lanes else None
        self._cap_key = key

    def _add_out_block(self):
        host = torch.empty(self._out_bytes, dtype=torch.uint8, pin_memory=True)
        host.numpy()[:] = 0
        self._out_blocks.append((host, host if self._zc_out else torch.zeros(self._out_bytes, dtype=torch.uint8, device=self.device)))

    def _bind_out_block(self, slot: int):
        """Make output block `slot` the one the next launch writes and `result` / `winner` / `winner_steps` view."""
        Bn, geom, ooffs = self.n_scen, self.geom, self._out_offs
        self._out_host, self._out_dev = self._out_blocks[slot]
        self.out_slot = slot
        ov = self._out_host.numpy()
        self.winner = ov[:Bn * 12 * geom.n_pad_max * 8].view(np.float64).reshape(Bn, 12, geom.n_pad_max)
        self.result = ov[ooffs["result"]:ooffs["result"] + Bn * nat.RESULT_BYTES].view(np.dtype(nat.RESULT_DTYPE))
        self.result_i32 = ov[ooffs["result"]:ooffs["result"] + Bn * nat.RESULT_BYTES].view(np.int32)     # (the planner's fast look-up of scenario 0's indices)
        self.winner_steps = ov[ooffs["winner_steps"]:ooffs["winner_steps"] + Bn * 4].view(np.int32)
        op = self._out_dev.data_ptr()
        self.B.result = op + ooffs["result"]
        self.B.winner, self.B.winner_steps = op + ooffs["winner"], op + ooffs["winner_steps"]
        if self.masked_winner:
            o = ooffs["winner_masked"]
            self.winner_masked = ov[o:o + Bn * 12 * geom.n_pad_max * 8].view(np.float64).reshape(Bn, 12, geom.n_pad_max)
            o = ooffs["winner_masked_steps"]
            self.winner_masked_steps = ov[o:o + Bn * 4].view(np.int32)
            self.B.winner_masked, self.B.winner_masked_steps = op + ooffs["winner_masked"], op + ooffs["winner_masked_steps"]

    def out_views(self, slot: int):
        """(result records, winner blocks, winner step counts) as NumPy views of output block `slot`'s pinned host copy."""
        Bn, geom, ooffs = self.n_scen, self.geom, self._out_offs
        ov = self._out_blocks[slot][0].numpy()
        return (ov[ooffs["result"]:ooffs["result"] + Bn * nat.RESULT_BYTES].view(np.dtype(nat.RESULT_DTYPE)),
                ov[:Bn * 12 * geom.n_pad_max * 8].view(np.float64).reshape(Bn, 12, geom.n_pad_max),
                ov[ooffs["winner_steps"]:ooffs["winner_steps"] + Bn * 4].view(np.int32))

    def enable_out_slots(self, n: int = 2):
        """A second output block (result records + winner trajectories), so that step i's results can leave the device on a
        side stream while step i + 1 already writes the other block (BatchPlanner.launch_pipelined)."""
        while len(self._out_blocks) < n:
            self._add_out_block()
        self._graphs.clear()

    def select_out_slot(self, slot: int):
        if slot != self.out_slot:
            self._bind_out_block(slot)

    def workspace_sizes(self) -> nat.FrenetSizes:
        """Byte sizes of every buffer of this shape as the C ABI reports them (frenet_workspace_bytes)."""
        out = nat.FrenetSizes()
        nat.check(nat.lib.frenet_workspace_bytes(C.byref(self.L), int(self.n_obs), int(self.n_obs_steps), C.byref(out)))
        return out

    def launch_shape(self, path: "PathHandle | None" = None, collision: int = 0) -> dict:
        """How the evaluate kernel launches for the configured shape (frenet_k2_launch_shape)."""
        rows, parts, grid, smem = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
        nat.check(nat.lib.frenet_k2_launch_shape(C.byref(self.L), C.byref(path.desc) if path is not None else None, int(self.B.n_obs),
                                                 int(self.B.n_obs_steps), int(collision), C.byref(rows), C.byref(parts), C.byref(grid),
                                                 C.byref(smem)))
        return {"rows_per_cta": rows.value, "ctas_per_row": parts.value, "grid": grid.value, "shared_bytes": smem.value}

    def rows_per_cta(self, with_xy: bool = True, collision: int = 0) -> int:
        path = next(iter(self._paths.values()))[1] if (with_xy and self._paths) else None
        return self.launch_shape(path, collision)["rows_per_cta"]

    def hbm_bytes(self) -> int:
        ts = [self.coef_lon, self.coef_lat, self.row_S, self.row_flags, self.lon, self.lat, self.cost, self.cand_flags,
              self.curv_ok, self.coll_free, self.first_blocker, self.road_ok, self._in_dev, self._out_dev, self._geom_dev] + [
            t for t in (self.obs_xy_t, self.obs_bounds) if t is not None]
        return sum(t.numel() * t.element_size() for t in ts)

    @staticmethod
    def bytes_per_scenario(geom: LatticeGeometry, n_v_cap: int) -> int:
        row = n_v_cap * geom.sum_pad
        cand = row * geom.n_d
        C_ = n_v_cap * geom.n_t * geom.n_d
        return 8 * (4 * row + 6 * cand) + C_ * (6 * 8 + 4 * 8 + 3 + 4) + n_v_cap * geom.n_t * (6 * 8 + 8 + 4)

    def set_relative_s(self, on: bool):
        """Lattice transform at `path_origin[b, 0] + (s - path_origin[b, 1])` = projection + (s - s0[0]) (frenet_to_glob_planner of the Duckietown driver,
        lib/tests/test_mapper_planner_obstacles.py:25-37) instead of at s.  `path_origin` is part of the packed inputs."""
        on = bool(on)
        if on != self.relative_s:
            self.relative_s = on
            self._graphs.clear()
            if self.geom is not None:
                self.B.path_origin = self._in_dev.data_ptr() + self._in_offs["path_origin"] if on else None

    def set_per_scenario_lanes(self, on: bool):
        """Batched lane runs: the parabola path and the lane boundaries are read per scenario from `lane_params` ([B][4]: a, b, c, ds)
        and `road_params` ([B][8]: right h, k, a, given; left h, k, a, given) of the packed inputs instead of the shared descriptors."""
        on = bool(on)
        if on != self.per_scenario_lanes:
            self.per_scenario_lanes = on
            self._graphs.clear()
            if self.geom is not None:
                ip = self._in_dev.data_ptr()
                self.B.lane_params = ip + self._in_offs["lane_params"] if on else None
                self.B.road_params = ip + self._in_offs["road_params"] if on else None

    @_on_device
    def offroad(self, right_fit=None, left_fit=None):
        """road_ok for every candidate against the lane boundary fits (a, b, c) of y = a x^2 + b x + c; None = that boundary
        was not detected (check_paths + check_offroad, lib/tests/test_mapper_planner_obstacles.py:62-70, 112-124)."""
        if not self.has_xy:
            raise RuntimeError("offroad(): the lattice has no Cartesian samples yet (run K3 first)")
        road = nat.FrenetRoad()
        for name, fit in (("right", right_fit), ("left", left_fit)):
            if fit is None:
                continue
            a, b, c = (float(v) for v in fit)
            vals = (-b / (2 * a), -(b ** 2 - 4 * a * c) / (4 * a), a)   # get_vertex_and_focus_distance (:105-109)
            arr = getattr(road, name)
            for i in range(3):
                arr[i] = vals[i]
            setattr(road, "has_" + name, 1)
        nat.check(nat.lib.frenet_k3_offroad(C.byref(self.L), C.byref(road), C.byref(self.B),
                                            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    # -- path handles --------------------------------------------------------------------------
    def path_handle(self, trajectory) -> PathHandle:
        key = id(trajectory)
        h = self._paths.get(key)
        if h is None or h[0] is not trajectory:
            h = (trajectory, make_path_handle(trajectory, self.device))
            self._paths[key] = h
        return h[1]

    # -- launches ------------------------------------------------------------------------------
    @_on_device
    def upload(self, record_event: bool = True):
        """One H2D copy of the packed per-call inputs (state, scalars, V axes, targets, obstacles).  An event marks the
        end of the copy so that the host may refill the pinned block for the NEXT call while this call's kernels run
        (`wait_staging_free`)."""
        if not self._zc_in:
            self._in_dev.copy_(self._in_host, non_blocking=True)
        if record_event:
            if not hasattr(self, "_h2d_done"):
                self._h2d_done = torch.cuda.Event()
            self._h2d_done.record(torch.cuda.current_stream(self.device))

    def wait_staging_free(self):
        """Block until the last `upload()` has finished reading the pinned staging block."""
        ev = getattr(self, "_h2d_done", None)
        if ev is not None:
            ev.synchronize()

    @_on_device
    def upload_region(self, *names):
        """H2D of individual input arrays of the packed block (e.g. "robot_pose", "obs_xy") without touching the rest -
        used when the planner state itself lives on the device between replans."""
        sizes = {"state": self.n_scen * 48, "scal": self.n_scen * 24, "axis_v": self.n_scen * self.n_v_cap * 8,
                 "target": self.n_scen * self.geom.n_t * 24, "obs_xy": self.n_scen * max(self.n_obs, 1) * 16,
                 "obs_sd": self.n_scen * max(self.n_obs, 1) * 16, "obs_speed": self.n_scen * max(self.n_obs, 1) * 8,
                 "robot_pose": self.n_scen * 24, "path_origin": self.n_scen * 16, "lane_params": self.n_scen * 32,
                 "road_params": self.n_scen * 64, "n_v": self.n_scen * 4, "obs_active": self.n_scen * max(self.n_obs, 1)}
        for name in (() if self._zc_in else names):
            o = self._in_offs[name]
            self._in_dev[o:o + sizes[name]].copy_(self._in_host[o:o + sizes[name]], non_blocking=True)

    # -- per-scenario follow mode on the device --------------------------------
    def enable_follow(self):
        """Device arrays of the batched follow fallback: per-scenario follow flag (wired into FrenetBuffers.follow), leading
        obstacle and re-plan mask."""
        if getattr(self, "follow_dev", None) is None or self.follow_dev.numel() != self.n_scen:
            self.follow_dev = torch.zeros(self.n_scen, dtype=torch.uint8, device=self.device)
            self.lead_dev = torch.full((self.n_scen,), -1, dtype=torch.int32, device=self.device)
            self.redo_dev = torch.zeros(self.n_scen, dtype=torch.uint8, device=self.device)
        self.B.follow = self.follow_dev.data_ptr()
        self._graphs.clear()

    def set_scenario_mask(self, mask: "torch.Tensor | None"):
        """Restrict the next launches to the scenarios with mask[b] != 0 (device uint8 [n_scen]); None = all."""
        self.B.scen_mask = mask.data_ptr() if mask is not None else None

    @_on_device
    def follow_targets(self, path: "PathHandle", estimate: torch.Tensor, delta_t: float, d_target: float, only: "torch.Tensor | None" = None):
        """frenet_follow_targets: target triples of the scenarios in follow mode from their leading obstacle's time law (the uploaded
        `obs_sd` / `obs_speed`), the robot's projection `estimate` (device [n_scen]) and the staged state."""
        ip = self._in_dev.data_ptr()
        n = self.n_scen * max(self.n_obs, 1)
        sd = ip + self._in_offs["obs_sd"]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib.frenet_follow_targets(C.byref(path.desc), C.byref(self.L), C.byref(self.B), self.lead_dev.data_ptr(), sd, sd + n * 8,
                                                ip + self._in_offs["obs_speed"], int(self.n_obs), estimate.data_ptr(), float(delta_t),
                                                float(d_target), only.data_ptr() if only is not None else None, C.c_void_p(stream)))

    @_on_device
    def follow_select(self, d_d_s: float, num_sample: int):
        """frenet_follow_select: scenarios without a survivor switch to follow mode behind their first blocker and get redo = 1."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib.frenet_follow_select(C.byref(self.L), C.byref(self.B), float(d_d_s), int(num_sample), self.lead_dev.data_ptr(),
                                               self.follow_dev.data_ptr(), self.redo_dev.data_ptr(), C.c_void_p(stream)))

    @_on_device
    def optimal_targets(self, spec, delta_t: float) -> torch.Tensor:
        """frenet_optimal_targets: the Optimal variant's sampled follow / merge / stop targets evaluated on the device into the
        target triples; spec [n_scen][16] (see the header).  Returns the device status array."""
        sp = torch.as_tensor(np.ascontiguousarray(spec, dtype=np.float64).reshape(self.n_scen, 16)).to(self.device)
        status = torch.zeros(self.n_scen, dtype=torch.int32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib.frenet_optimal_targets(C.byref(self.L), C.byref(self.B), sp.data_ptr(), float(delta_t), status.data_ptr(),
                                                 C.c_void_p(stream)))
        self._opt_spec = sp
        return status

    @_on_device
    def advance_state(self, time, lookup_period: float, d_d_s: float, num_sample: int, follow: bool = False):
        """Device-side replan bookkeeping (frenet_advance_state): next state = gathered winner at `time`, new V axes (the target
        offsets of si_interval for scenarios in follow mode: `follow`, or per scenario through FrenetBuffers.follow)."""
        if not hasattr(self, "_time_dev") or self._time_dev.numel() != self.n_scen:
            self._time_dev = torch.empty(self.n_scen, dtype=torch.float64, device=self.device)
            self._adv_status = torch.zeros(self.n_scen, dtype=torch.int32, device=self.device)
        if np.ndim(time) == 0:
            self._time_dev.fill_(float(time))
        else:
            self._time_dev.copy_(torch.as_tensor(np.ascontiguousarray(time, dtype=np.float64)), non_blocking=True)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib.frenet_advance_state(C.byref(self.L), C.byref(self.B), self._time_dev.data_ptr(), 1, float(lookup_period),
                                               float(d_d_s), int(num_sample), 1 if follow else 0, self._adv_status.data_ptr(),
                                               C.c_void_p(stream)))

    def device_state(self):
        """(state [B][6], scal [B][3], n_v [B]) as they currently are ON THE DEVICE."""
        o, B = self._in_offs, self.n_scen
        st = self._in_dev[o["state"]:o["state"] + B * 48].cpu().numpy().view(np.float64).reshape(B, 6)
        sc = self._in_dev[o["scal"]:o["scal"] + B * 24].cpu().numpy().view(np.float64).reshape(B, 3)
        nv = self._in_dev[o["n_v"]:o["n_v"] + B * 4].cpu().numpy().view(np.int32)
        return st, sc, nv

    @_on_device
    def launch(self, params: nat.FrenetParams, path: PathHandle | None, stages: int, use_mask: int = 0,
               gather_sel: int = nat.SEL_CTOT, n_obs: int | None = None):
        """Enqueue the selected pipeline stages on torch's current stream (asynchronous)."""
        if n_obs is not None:
            if n_obs > self.n_obs:
                raise ValueError(f"{n_obs} obstacles exceed the configured capacity {self.n_obs}")
            self.B.n_obs = int(n_obs)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        pp = C.byref(path.desc) if path is not None else None
        nat.check(nat.lib.frenet_plan(C.byref(self.L), C.byref(params), pp, C.byref(self.B), int(stages), int(use_mask),
                                      int(gather_sel), C.c_void_p(stream)))
        self._pending = True
        if stages & nat.STAGE_K2:
            self.generation += 1
            self.has_xy = False
            self.has_collision = False
        if stages & nat.STAGE_K3:
            self.has_xy = True
        if stages & nat.STAGE_K4:
            self.has_collision = True

    @_on_device
    def place_obstacles(self, path: PathHandle):
        """Frenet -> Cartesian of the uploaded obstacle table `obs_sd` ([2][B][O]: arc lengths, lateral offsets)
        straight into the device `obs_xy` table ([B][O][2]) that K4 reads - the device side of
        `sensor.step_obstacle` + `ObstacleTrajectory.compute_pt` (lib/sensor/proximity_sensor.py:82-89,
        lib/sensor/dynamic_obstacle.py:46-56)."""
        ip = self._in_dev.data_ptr()
        n = self.n_scen * max(self.n_obs, 1)
        sd = ip + self._in_offs["obs_sd"]
        xy = ip + self._in_offs["obs_xy"]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib.frenet_points_to_cartesian(C.byref(path.desc), sd, sd + n * 8, n, xy, xy + 8, 2, C.c_void_p(stream)))

    @_on_device
    def sensor_gate(self, rng: float, aperture: float):
        """Device-side `ProximitySensor.sense` for the whole batch: fills the device `obs_active` mask from the uploaded
        `robot_pose` and the device `obs_xy` table (lib/sensor/proximity_sensor.py:41-75)."""
        ip = self._in_dev.data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib.frenet_sensor_gate(ip + self._in_offs["robot_pose"], ip + self._in_offs["obs_xy"], self.n_scen, self.n_obs,
                                             float(rng), float(aperture), ip + self._in_offs["obs_active"], C.c_void_p(stream)))

    def device_obs_active(self) -> np.ndarray:
        o = self._in_offs["obs_active"]
        return self._in_dev[o:o + self.n_scen * max(self.n_obs, 1)].cpu().numpy().reshape(self.n_scen, max(self.n_obs, 1))

    @_on_device
    def predict_obstacles(self, path: PathHandle, period: float):
        """EXTENSION: fill the predicted table from the uploaded `obs_sd` / `obs_speed` and each scenario's t_initial with the
        reference's moving-obstacle time law (lib/sensor/dynamic_obstacle.py:43-56), one entry per sample period."""
        if self.obs_xy_t is None:
            raise RuntimeError("engine was configured without a predicted obstacle table (n_obs_steps = 0)")
        ip = self._in_dev.data_ptr()
        n = self.n_scen * self.n_obs
        sd = ip + self._in_offs["obs_sd"]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib.frenet_predict_obstacles(C.byref(path.desc), sd, sd + n * 8, ip + self._in_offs["obs_speed"],
                                                   ip + self._in_offs["scal"] + 16, 3, self.n_scen, self.n_obs, self.n_obs_steps,
                                                   float(period), self.obs_xy_t.data_ptr(), C.c_void_p(stream)))

    @_on_device
    def download(self, sync: bool = True):
        """One D2H copy of the packed per-call outputs (result records, winner trajectories)."""
        if not self._zc_out:
            self._out_host.copy_(self._out_dev, non_blocking=True)
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
            self._pending = False

    def host_block_free(self):
        """zero_copy engines: before the host rewrites the pinned input block, wait for launches that may still be reading it
        (only a launch that nobody waited for leaves the flag set - the graph paths end with a synchronize)."""
        if self._zc_in and self._pending:
            torch.cuda.current_stream(self.device).synchronize()
            self._pending = False

    def run(self, params, path, stages, use_mask=0, gather_sel=nat.SEL_CTOT, n_obs=None):
        """upload -> pipeline -> download, synchronous."""
        self.upload()
        self.launch(params, path, stages, use_mask, gather_sel, n_obs)
        self.download(sync=True)

    @_on_device
    def run_graph(self, params, path, stages, use_mask=0, gather_sel=nat.SEL_CTOT, n_obs=None):
        """`run` replayed from a CUDA graph (H2D copy, kernels, D2H copy as one launch): the launch-bound single-replan
        path.  The graph bakes in the kernel arguments (parameters, sizes, pointers), so it is keyed by them and
        re-captured when any of them changes; the per-replan inputs travel through the pinned block, which the graph's
        copy node reads at replay time."""
        # (configure() clears the graphs, so the capacities need not be part of the key; the parameter block is keyed by its bytes:
        #  callers may mutate it in place)
        key = (bytes(params), id(path), stages, use_mask, gather_sel, self.B.n_obs if n_obs is None else n_obs, self.out_slot)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) > 16:
                self._graphs.clear()
            self.run(params, path, stages, use_mask, gather_sel, n_obs)      # warm-up outside capture (lazy module load, attributes)
            gen, has_xy, has_coll = self.generation, self.has_xy, self.has_collision
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.upload(record_event=False)
                self.launch(params, path, stages, use_mask, gather_sel, n_obs)
                self.download(sync=False)
            self.generation, self.has_xy, self.has_collision = gen, has_xy, has_coll
            self._graphs[key] = g
            self._pending = False
            return                      # the warm-up run already produced this call's results
        g.replay()
        # (the raw handle + one foreign call: building a torch Stream object just to wait on it costs ~4 us of a ~30 us replan)
        nat.check(nat.lib.frenet_stream_synchronize(torch._C._cuda_getCurrentRawStream(self._dev_index)))
        self._pending = False
        if stages & nat.STAGE_K2:
            self.generation += 1
            self.has_xy = bool(stages & nat.STAGE_K3)
            self.has_collision = bool(stages & nat.STAGE_K4)

    @_on_device
    def run_inline(self, params, path, stages, use_mask=0, gather_sel=nat.SEL_CTOT, n_obs=None, follow: bool = False):
        """ONE planner's replan as one kernel launch and nothing else (frenet_replan_inline): the library reads state / scalars /
        V axis (/ targets) from the pinned staging block at call time and hands them to the kernel as an ARGUMENT - no copy node,
        no graph - and returns when the stream has drained.  Falls back to `run_graph` when the shape exceeds the inline
        capacities.  The obstacle table is NOT part of the inline inputs: a launch that includes K4 tests the table that is on
        the device (the planner only does that with the obstacles its last check_paths uploaded)."""
        if self.n_scen != 1 or self.n_v_cap > nat.INLINE_MAX_V or (follow and self.geom.n_t > nat.INLINE_MAX_T) or self._zc_in:
            return self.run_graph(params, path, stages, use_mask, gather_sel, n_obs)
        if n_obs is not None:
            if n_obs > self.n_obs:
                raise ValueError(f"{n_obs} obstacles exceed the configured capacity {self.n_obs}")
            self.B.n_obs = int(n_obs)
        hp, o = self._in_host.data_ptr(), self._in_offs
        stream = torch._C._cuda_getCurrentRawStream(self._dev_index)
        nat.check(nat.lib.frenet_replan_inline(C.byref(self.L), C.byref(params), C.byref(path.desc) if path is not None else None,
                                               C.byref(self.B), int(stages), int(use_mask), int(gather_sel), hp + o["state"], hp + o["scal"],
                                               hp + o["axis_v"], hp + o["n_v"], hp + o["target"] if follow else None, 1, stream))
        self._pending = False
        if stages & nat.STAGE_K2:
            self.generation += 1
            self.has_xy = bool(stages & nat.STAGE_K3)
            self.has_collision = bool(stages & nat.STAGE_K4)

    # -- views of device data ------------------------------------------------------------------
    def cand_offsets(self, n_v: int):
        """Per candidate in list order: offset of its [6][npad] block in a scenario's `lat` slab, offset of its row's
        [4][npad] block in the `lon` slab, padded length, step count."""
        g = self.geom
        iv = np.repeat(np.arange(n_v), g.n_t * g.n_d)
        iT = np.tile(np.repeat(np.arange(g.n_t), g.n_d), n_v)
        idd = np.tile(np.arange(g.n_d), n_v * g.n_t)
        within = iv.astype(np.int64) * g.sum_pad + g.t_offset[iT]
        npad = g.n_pad[iT].astype(np.int64)
        return (within * g.n_d + idd * npad) * 6, within * 4, npad, g.n_steps[iT]

    def fetch_scenario(self, b: int = 0, with_xy: bool | None = None, with_collision: bool | None = None) -> dict:
        """D2H of one scenario's full lattice state as NumPy arrays (raw padded layout + offsets)."""
        with_xy = self.has_xy if with_xy is None else with_xy
        with_collision = self.has_collision if with_collision is None else with_collision
        nv = int(self.n_v[b])
        out = dict(n_v=nv, lon=self.lon[b].cpu().numpy(), lat=self.lat[b].cpu().numpy(),
                   cost=self.cost[:, b].cpu().numpy(), cand_flags=self.cand_flags[b].cpu().numpy(),
                   coef_lon=self.coef_lon[b].cpu().numpy(), coef_lat=self.coef_lat[b].cpu().numpy(),
                   row_S=self.row_S[b].cpu().numpy(), row_flags=self.row_flags[b].cpu().numpy())
        out["has_xy"] = bool(with_xy)
        if with_collision:
            out["coll_free"] = self.coll_free[b].cpu().numpy()
            out["first_blocker"] = self.first_blocker[b].cpu().numpy()
        out["lat_off"], out["lon_off"], out["npad"], out["nsteps"] = self.cand_offsets(nv)
        return out

    def padded_fields(self, sc: dict) -> dict:
        """Scenario dict -> [C][Nmax] NaN-padded arrays in the golden-fixture layout (tests, lazy views)."""
        g = self.geom
        C_ = sc["n_v"] * g.n_t * g.n_d
        nmax = int(g.n_steps.max())
        k = np.arange(nmax)
        n = sc["nsteps"]
        valid = k[None, :] < n[:, None]
        il = np.where(valid, sc["lat_off"][:, None] + k[None, :], 0)
        io = np.where(valid, sc["lon_off"][:, None] + k[None, :], 0)
        fstride = np.where(valid, sc["npad"][:, None], 0)
        res = {"nsteps": n.astype(np.int32)}
        names = ("d", "dot_d", "ddot_d", "dddot_d") + (("x", "y") if sc.get("has_xy") else ())
        for j, name in enumerate(names):
            res[name] = np.where(valid, sc["lat"][il + j * fstride], np.nan)
        for j, name in enumerate(("s", "dot_s", "ddot_s", "dddot_s")):
            res[name] = np.where(valid, sc["lon"][io + j * fstride], np.nan)
        for j, name in enumerate(("cd", "cv", "ct", "ctot")):
            res[name] = sc["cost"][j][:C_]
        res["valid"] = (sc["cand_flags"][:C_] & nat.CAND_VALID) != 0
        res["kinematic_ok"] = (sc["cand_flags"][:C_] & nat.CAND_KINEMATIC_OK) != 0
        if "coll_free" in sc:
            res["coll_free"] = sc["coll_free"][:C_] != 0
            res["first_blocker"] = sc["first_blocker"][:C_]
        return res

    @_on_device
    def points_to_cartesian(self, path: PathHandle, s, d):
        """Point-wise Frenet -> Cartesian on the device; s, d array-likes of equal length; returns (x, y) NumPy."""
        s = torch.as_tensor(np.ascontiguousarray(s, dtype=np.float64)).to(self.device)
        d = torch.as_tensor(np.ascontiguousarray(d, dtype=np.float64)).to(self.device)
        x = torch.empty_like(s)
        y = torch.empty_like(s)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(nat.lib.frenet_points_to_cartesian(C.byref(path.desc), s.data_ptr(), d.data_ptr(), s.numel(), x.data_ptr(),
                                                     y.data_ptr(), 1, C.c_void_p(stream)))
        return x.cpu().numpy(), y.cpu().numpy()

    @_on_device
    def polyline_collision(self, xs, ys, obstacles_xy, radius_sq: float):
        """Generic K4: `xs`, `ys` are lists of per-path coordinate sequences; returns (coll_free bool[n], first_blocker int[n])."""
        n = len(xs)
        offsets = np.zeros(n + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(v) for v in xs])
        x = np.concatenate([np.asarray(v, dtype=np.float64) for v in xs]) if n else np.zeros(0)
        y = np.concatenate([np.asarray(v, dtype=np.float64) for v in ys]) if n else np.zeros(0)
        obs = np.ascontiguousarray(obstacles_xy, dtype=np.float64).reshape(-1, 2)
        dev = self.device
        tx, ty, to, tb = (torch.from_numpy(a).to(dev) for a in (x, y, offsets, obs))
        free = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
        first = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        nat.check(nat.lib.frenet_polyline_collision(tx.data_ptr(), ty.data_ptr(), to.data_ptr(), n, tb.data_ptr(), obs.shape[0],
                                                    float(radius_sq), free.data_ptr(), first.data_ptr(), C.c_void_p(stream)))
        return free.cpu().numpy()[:n] != 0, first.cpu().numpy()[:n]
