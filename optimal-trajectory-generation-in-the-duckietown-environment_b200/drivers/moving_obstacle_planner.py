"""Moving-obstacle experiment (lib/tests/test_moving_obstacle_planner.py): same function names and signatures."""
from __future__ import annotations

import numpy as np

from ..sensor import ProximitySensor, MovingObstacle
from .common import (robot_radius, compute_ortogonal_vect, frenet_to_glob, collision_filter, single_path_collision,
                     best_path, LeadingObstacles)
from .obstacle_planner import seeded_obstacles
from .simulation import SimulationResult, default_elements, frenet_start


def check_paths(frenet_paths, sensor, rpose):
    """(surviving candidates, one blocking obstacle per blocked candidate) - test_moving_obstacle_planner.py:26-39."""
    kept, sensed, first = collision_filter(frenet_paths, sensor, rpose, want_blockers=True)
    return kept, LeadingObstacles(sensed, first)


def check_collisions(path, obstacles):
    """(collision_free, first blocking obstacle or None) - test_moving_obstacle_planner.py:42-52."""
    return single_path_collision(path, obstacles)


def simulate(t_vect, trajectory, robot, transformer, controller, planner, sensor, dt=0.1, hook=None) -> SimulationResult:
    """The closed loop of test_moving_obstacle_planner.py:60-124, including the fall-back to follow mode behind the
    first blocking obstacle when every candidate collides (:92-97)."""
    res = SimulationResult(len(t_vect))
    robot_p = robot.p
    s0, d0 = frenet_start(trajectory, transformer, robot_p)
    planner.initialize(t0=t_vect[0], p0=d0, s0=s0)
    for i, t in enumerate(t_vect):
        est_pt = transformer.estimatePosition(trajectory, robot_p)
        robot_fpose = transformer.transform(robot_p)
        pos_s, pos_d = planner.replanner(t)
        paths = frenet_to_glob(planner, trajectory, planner.paths)
        sensor.step_obstacle(t)
        planner.paths, leading = check_paths(paths, sensor, robot_p)
        if len(planner.paths) == 0:
            s_target = leading[0].trajectory
            s_target.set_transformer(transformer)
            pos_s, pos_d = planner.replanner(t, s_target=s_target)
            frenet_to_glob(planner, trajectory, planner.paths)
            res.follow_steps.append(i)
        planner.opt_path_tot = best_path(planner.paths)
        if hook is not None:
            hook(i, planner)
        error = np.array([0, pos_d[0]]) - robot_fpose[0:2]
        derror = np.array([pos_s[1], pos_d[1]])
        u = controller.compute(robot_fpose, error, derror, trajectory.compute_curvature(est_pt))
        robot_p, _ = robot.step(u, dt)
        res.record(i, robot_p, robot_fpose, u, error, derror,
                   trajectory.compute_pt(pos_s[0]) + compute_ortogonal_vect(trajectory, pos_s[0]) * pos_d[0])
    return res


def test_planner_moving_obstacle(t_vect=None, trajectory=None, **overrides) -> SimulationResult:
    """Headless `tester.py -t planner_moving_obstacles` on the circle configuration (:127-173)."""
    el = default_elements(t_vect, trajectory, **overrides)
    pts = seeded_obstacles(el.t_vect, el.trajectory)
    obstacles = [MovingObstacle(pts[:, i], el.trajectory) for i in range(pts.shape[1])]
    sensor = ProximitySensor(3.5, np.pi / 6)
    sensor.attach(el.robot)
    sensor.set_obstacle(obstacles)
    return simulate(el.t_vect, el.trajectory, el.robot, el.transformer, el.controller, el.planner, sensor)
