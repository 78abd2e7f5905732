"""What the three experiments share: default scene (config/config_circle2D.json decoded by hand, SURVEY.md section 2
row 22; lib/tests/config.py:74-140) and the per-step log (lib/logger/simulation_logger.py rows)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from ..planner import TrajectoryPlannerV1, TrajectoryPlannerParams
from ..sim import FrenetGNTransformOld, FrenetIOLController, Unicycle
from ..trajectory import CircleTrajectory2D


@dataclass
class Elements:
    t_vect: np.ndarray
    trajectory: object
    robot: Unicycle
    transformer: FrenetGNTransformOld
    controller: FrenetIOLController
    planner: object


def default_elements(t_vect=None, trajectory=None, robot_pose=(0.0, 0.0, 0.0), planner=None) -> Elements:
    t_vect = np.arange(0.1, 50, 0.1) if t_vect is None else t_vect
    robot = Unicycle()
    robot.set_initial_pose(np.array(robot_pose, dtype=np.float64))
    return Elements(t_vect, CircleTrajectory2D(8, 5, 2) if trajectory is None else trajectory, robot,
                    FrenetGNTransformOld(0.5, 10), FrenetIOLController(.5, 0.0, 5.0, 0.0, 0.0),
                    TrajectoryPlannerV1(TrajectoryPlannerParams()) if planner is None else planner)


def frenet_start(trajectory, transformer, robot_p):
    """Initial planner state from the robot's projection (test_planner.py:33-37)."""
    transformer.estimatePosition(trajectory, robot_p)
    f = transformer.transform(robot_p)
    return (f[0], 0.0, 0.0), (f[1], 0.0, 0.0)


@dataclass
class SimulationResult:
    n: int
    follow_steps: list = field(default_factory=list)

    def __post_init__(self):
        n = self.n
        self.robot_pose = np.zeros((n, 3))
        self.robot_frenet_pose = np.zeros((n, 3))
        self.control = np.zeros((n, 2))
        self.error = np.zeros((n, 2))
        self.derror = np.zeros((n, 2))
        self.target = np.zeros((n, 2))

    def record(self, i, pose, fpose, u, error, derror, target):
        self.robot_pose[i], self.robot_frenet_pose[i], self.control[i] = pose, fpose, u
        self.error[i], self.derror[i], self.target[i] = error, derror, target
