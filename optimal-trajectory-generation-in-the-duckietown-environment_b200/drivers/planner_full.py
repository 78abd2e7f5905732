"""Obstacle-free experiment (lib/tests/test_planner.py:19-105)."""
from __future__ import annotations

import numpy as np

from .common import compute_ortogonal_vect
from .simulation import SimulationResult, default_elements, frenet_start


def simulate(t_vect, trajectory, robot, transformer, controller, planner, dt=0.1, hook=None) -> SimulationResult:
    res = SimulationResult(len(t_vect))
    robot_p = robot.p
    s0, d0 = frenet_start(trajectory, transformer, robot_p)
    planner.initialize(t0=t_vect[0], p0=d0, s0=s0)
    for i, t in enumerate(t_vect):
        est_pt = transformer.estimatePosition(trajectory, robot_p)
        robot_fpose = transformer.transform(robot_p)
        pos_s, pos_d = planner.replanner(t)
        if hook is not None:
            hook(i, planner)
        error = np.array([0, pos_d[0]]) - robot_fpose[0:2]
        derror = np.array([pos_s[1], pos_d[1]])
        u = controller.compute(robot_fpose, error, derror, trajectory.compute_curvature(est_pt))
        robot_p, _ = robot.step(u, dt)
        res.record(i, robot_p, robot_fpose, u, error, derror,
                   trajectory.compute_pt(pos_s[0]) + compute_ortogonal_vect(trajectory, pos_s[0]) * pos_d[0])
    return res


def test_planner_full(t_vect=None, trajectory=None, **overrides) -> SimulationResult:
    """Headless `tester.py -t planner_full` (no plots)."""
    el = default_elements(t_vect, trajectory, **overrides)
    return simulate(el.t_vect, el.trajectory, el.robot, el.transformer, el.controller, el.planner)
