"""Static-obstacle experiment (lib/tests/test_obstacle_planner.py): same function names and signatures."""
from __future__ import annotations

import numpy as np

from ..sensor import ProximitySensor, StaticObstacle
from .common import (robot_radius, compute_ortogonal_vect, frenet_to_glob, collision_filter, single_path_collision,
                     best_path)
from .simulation import SimulationResult, default_elements, frenet_start


def check_paths(frenet_paths, sensor, rpose):
    """Candidates that do not touch a sensed obstacle, order preserved (test_obstacle_planner.py:27-38)."""
    kept, _, _ = collision_filter(frenet_paths, sensor, rpose)
    return kept


def check_collisions(path, obstacles) -> bool:
    """True when `path` keeps more than robot_radius away from every obstacle (test_obstacle_planner.py:41-51)."""
    return single_path_collision(path, obstacles)[0]


def seeded_obstacles(t_vect, trajectory):
    """The ten obstacles of the reference run (test_obstacle_planner.py:142-157): np.random.seed(5), positions on the
    path at ten random arc lengths, N(0, 2) noise on y."""
    np.random.seed(5)
    t_samples = np.random.choice(t_vect, 10)
    pts = np.array([trajectory.compute_pt(t) for t in t_samples]).T
    pts[1, :] += np.random.normal(0, 2., pts.shape[1])
    return pts


def simulate(t_vect, trajectory, robot, transformer, controller, planner, sensor, dt=0.1, hook=None) -> SimulationResult:
    """The closed loop of test_obstacle_planner.py:58-116: replan, transform, filter, re-pick, track."""
    res = SimulationResult(len(t_vect))
    robot_p = robot.p
    s0, d0 = frenet_start(trajectory, transformer, robot_p)
    planner.initialize(t0=t_vect[0], p0=d0, s0=s0)
    for i, t in enumerate(t_vect):
        est_pt = transformer.estimatePosition(trajectory, robot_p)
        robot_fpose = transformer.transform(robot_p)
        pos_s, pos_d = planner.replanner(t)
        paths = frenet_to_glob(planner, trajectory, planner.paths)
        planner.paths = check_paths(paths, sensor, robot_p)
        planner.opt_path_tot = best_path(planner.paths)          # ValueError when every candidate collides
        if hook is not None:
            hook(i, planner)
        error = np.array([0, pos_d[0]]) - robot_fpose[0:2]
        derror = np.array([pos_s[1], pos_d[1]])
        u = controller.compute(robot_fpose, error, derror, trajectory.compute_curvature(est_pt))
        robot_p, _ = robot.step(u, dt)
        res.record(i, robot_p, robot_fpose, u, error, derror,
                   trajectory.compute_pt(pos_s[0]) + compute_ortogonal_vect(trajectory, pos_s[0]) * pos_d[0])
    return res


def test_planner_obstacle(t_vect=None, trajectory=None, **overrides) -> SimulationResult:
    """Headless `tester.py -t planner_obstacles` on the circle configuration (test_obstacle_planner.py:119-166)."""
    el = default_elements(t_vect, trajectory, **overrides)
    pts = seeded_obstacles(el.t_vect, el.trajectory)
    obstacles = [StaticObstacle(pts[:, i]) for i in range(pts.shape[1])]
    sensor = ProximitySensor(3.5, np.pi / 6)
    sensor.attach(el.robot)
    sensor.set_obstacle(obstacles)
    return simulate(el.t_vect, el.trajectory, el.robot, el.transformer, el.controller, el.planner, sensor)
