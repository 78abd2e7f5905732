"""Obstacle types (lib/sensor/obstacle.py, static_obstacle.py:10-17, dynamic_obstacle.py:13-83).

Host-side producers of the obstacle table the collision kernel (K4) consumes: one (x, y) per obstacle.
"""
from __future__ import annotations

import numpy as np


class Obstacle:
    position = None


class StaticObstacle(Obstacle):
    def __init__(self, p: np.ndarray):
        self.position = p


class ObstacleTrajectory:
    """An obstacle riding the reference path: arc length `time_offset + (t * speed) % 50`, lateral `offset`."""

    def __init__(self, o: float, s: float, t0: float, t):
        self.offset = o
        self.speed = s
        self.time_offset = t0
        self.trajectory = t
        self.transformer = None      # the reference builds a default FrenetGNTransform here; drivers set their own

    def time_law(self, t):
        return self.time_offset + t * self.speed % 50

    def compute_orientation(self, t: float):
        g = self.trajectory.compute_first_derivative(self.time_law(t))
        return np.arctan2(g[1], g[0])

    def compute_ortogonal_vect(self, t: float) -> np.ndarray:
        th = self.compute_orientation(t)
        return np.array([-np.sin(th), np.cos(th)])

    def compute_pt(self, t):
        return self.trajectory.compute_pt(self.time_law(t)) + self.compute_ortogonal_vect(t) * self.offset

    # NB: like the reference (dynamic_obstacle.py:58-68) the derivatives are those of the underlying path at `t`
    def compute_first_derivative(self, t):
        return self.trajectory.compute_first_derivative(t)

    def compute_second_derivative(self, t):
        return self.trajectory.compute_second_derivative(t)

    def compute_third_derivative(self, t):
        return self.trajectory.compute_third_derivative(t)

    def compute_curvature(self, t):
        return self.trajectory.compute_curvature(t)

    def compute_s(self, t):
        return self.transformer.transform(self.compute_pt(t))[0]

    def compute_ds(self, t):
        return self.transformer.transform(self.compute_first_derivative(t))[0]

    def compute_dds(self, t):
        return self.transformer.transform(self.compute_second_derivative(t))[0]

    def compute_ddds(self, t):
        return self.transformer.transform(self.compute_third_derivative(t))[0]

    def set_transformer(self, transformer):
        self.transformer = transformer


class MovingObstacle(Obstacle):
    """Draws its lateral offset, speed and phase from NumPy's global RNG in the reference's order
    (randint(-1, 1), random(), randint(5, 10) - dynamic_obstacle.py:18-19) so seeded runs match."""

    def __init__(self, p: np.ndarray, t):
        self.start_position = p
        self.trajectory = ObstacleTrajectory(o=np.random.randint(-1, 1), s=np.random.random(), t0=np.random.randint(5, 10), t=t)
        self.step(0)

    def step(self, time: float):
        self.position = self.trajectory.compute_pt(time)
        self.orientation = self.trajectory.compute_orientation(time)
        self.pose = np.concatenate([self.position, self.orientation[None]], axis=0)

    def reset(self):
        self.position = self.start_position
