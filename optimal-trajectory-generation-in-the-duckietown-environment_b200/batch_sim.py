"""Closed-loop simulation of many agents with everything per-step resident on the device.

One `step(time)` is one iteration of the reference's experiment loops (lib/tests/test_planner.py:39-76,
lib/tests/test_obstacle_planner.py:75-114, lib/tests/test_moving_obstacle_planner.py:76-122) for every agent:

    transformer.estimatePosition / transform      -> frenet_sim_project          (robot pose -> Frenet pose, curvature)
    planner.replanner(t)                          -> frenet_advance_state + K1 + fused K2/K3/K4 + K5 + gather
    sensor.step_obstacle / sensor.sense           -> obstacle table upload + frenet_sensor_gate
    controller.compute + robot.step               -> frenet_sim_control_step

Only the moving obstacles' positions (when they move) go host -> device per step; nothing has to come back.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _native as nat
from .batch import BatchPlanner


class BatchSimulator:
    def __init__(self, params, trajectory, n_agents: int, n_obs: int = 0, variant: str = "v1", device=None,
                 controller=(0.5, 0.0, 5.0), gn=(0.5, 10, 0.01), dt: float = 0.1, sensor=(3.5, math.pi / 6)):
        self.bp = BatchPlanner(params, trajectory, n_agents, n_obs, variant, device)
        self.n = int(n_agents)
        self.dt = float(dt)
        self.ctrl_b, self.kp_s, self.kp_d = (float(x) for x in controller)
        self.gn_guess, self.gn_iter, self.gn_damping = float(gn[0]), int(gn[1]), float(gn[2])
        self.sensor = sensor
        dev = self.bp.engine.device
        f64 = dict(dtype=torch.float64, device=dev)
        self.estimate = torch.full((self.n,), self.gn_guess, **f64)
        self.fpose = torch.zeros((self.n, 3), **f64)
        self.curvature = torch.zeros((self.n,), **f64)
        self.control = torch.zeros((self.n, 2), **f64)

    # device addresses inside the engine's packed input block
    def _ptr(self, name):
        e = self.bp.engine
        return e._in_dev.data_ptr() + e._in_offs[name]

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.bp.engine.device).cuda_stream)

    def _project(self):
        nat.check(nat.lib.frenet_sim_project(C.byref(self.bp.path.desc), self._ptr("robot_pose"), self.estimate.data_ptr(), self.gn_iter,
                                             self.gn_damping, self.fpose.data_ptr(), self.curvature.data_ptr(), self.n, self._stream()))

    def enable_follow_fallback(self, obs_time_offset, obs_offset, obs_speed, d_target: float = 2.0):
        """All-blocked agents re-plan in follow mode behind their first blocker, on the device (BatchPlanner.enable_follow_fallback)."""
        self.bp.enable_follow_fallback(obs_time_offset, obs_offset, obs_speed, d_target)

    def reset(self, robot_pose, t0: float, obs_xy=None, s0=None):
        """Place the agents and initialise their planners like the experiments do (test_planner.py:31-38): project the robot,
        start the lattice from (s, 0, 0), (d, 0, 0) of its Frenet pose (s0 [B]: start arc lengths given explicitly instead)."""
        e = self.bp.engine
        self.estimate.fill_(self.gn_guess)
        e.robot_pose[:] = np.asarray(robot_pose, dtype=np.float64).reshape(self.n, 3)
        e.upload_region("robot_pose")
        self._project()
        f = self.fpose.cpu().numpy()
        z = np.zeros(self.n)
        kw = {}
        if self.bp.n_obs:
            kw = dict(obs_xy=np.zeros((self.n, self.bp.n_obs, 2)) if obs_xy is None else obs_xy,
                      obs_active=np.zeros((self.n, self.bp.n_obs), dtype=np.uint8))     # initialize() does not look at obstacles
        e.robot_pose[:] = np.asarray(robot_pose, dtype=np.float64).reshape(self.n, 3)   # plan() re-uploads the whole block
        s_start = f[:, 0] if s0 is None else np.broadcast_to(np.asarray(s0, dtype=np.float64), (self.n,))
        return self.bp.plan(np.stack([f[:, 1], z, z], 1), np.stack([s_start, z, z], 1), np.full(self.n, float(t0)), **kw)

    def step(self, time: float, obs_xy=None, sync: bool = False):
        """One closed-loop step for every agent; returns the (pinned) result records if `sync`."""
        e, bp = self.bp.engine, self.bp
        self._project()
        if obs_xy is not None:
            e.obs_xy[:] = obs_xy
            e.upload_region("obs_xy")
        lookup = getattr(bp.p, "sampling_t", None) if bp.variant == "dt_obstacles" else bp.p.delta_t
        e.advance_state(time, lookup, bp.p.d_d_s, bp.p.num_sample)
        if bp.n_obs and self.sensor is not None:
            e.sensor_gate(float(self.sensor[0]), float(self.sensor[1]))
        bp.pipeline_with_follow(self.estimate)
        nat.check(nat.lib.frenet_sim_control_step(self.fpose.data_ptr(), self._ptr("state"), self.curvature.data_ptr(), self.ctrl_b,
                                                  self.kp_s, self.kp_d, self.dt, self._ptr("robot_pose"), self.control.data_ptr(),
                                                  self.n, self._stream()))
        if sync:
            e.download(sync=True)
            return e.result
        return None

    def robot_pose(self) -> np.ndarray:
        e = self.bp.engine
        o = e._in_offs["robot_pose"]
        return e._in_dev[o:o + self.n * 24].cpu().numpy().view(np.float64).reshape(self.n, 3)
