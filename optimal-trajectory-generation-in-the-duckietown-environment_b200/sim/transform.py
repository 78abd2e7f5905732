"""Projection of the robot onto the reference path (lib/transform/gn_transform_old.py:11-106)."""
from __future__ import annotations

import numpy as np


class FrenetGNTransformOld:
    """Damped Gauss-Newton search of the arc length whose path point is closest to the robot, and the
    SE(2) change of frame into the Frenet frame anchored there."""

    def __init__(self, initial_guess: float = 0.5, max_iter: int = 10):
        self.estimate = initial_guess
        self.max_iter = max_iter
        self.robot_pose = None
        self.damping = 0.01

    def _gn_step(self, s: float, target: np.ndarray) -> float:
        residual = self.trajectory.compute_pt(s) - target
        jac = self.trajectory.compute_first_derivative(s)
        return -np.matmul(jac.T, residual) / (np.matmul(jac.T, jac) + self.damping)

    def estimatePosition(self, trajectory, p: np.ndarray) -> float:
        self.trajectory = trajectory
        self.robot_pose = p
        s = self.estimate
        for _ in range(self.max_iter):
            s += self._gn_step(s, p[0:2])
        self.estimate = s
        self.t_vect = trajectory.compute_pt(s)
        grad = trajectory.compute_first_derivative(s)
        self.t_r = np.arctan2(grad[1], grad[0])
        c, sn = np.cos(self.t_r), np.sin(self.t_r)
        self.r_mat = np.array([[c, -sn], [sn, c]])
        return self.estimate

    def transform(self, p: np.ndarray) -> np.ndarray:
        """Global pose (3,) or point (2,) -> Frenet frame."""
        Rt = self.r_mat.T
        if p.shape[0] == 3:
            return np.append(np.matmul(Rt, p[0:2]) - np.matmul(Rt, self.t_vect), p[2] - self.t_r)
        if p.shape[0] == 2:
            return np.matmul(Rt, p) - np.matmul(Rt, self.t_vect)
        raise ValueError('input array has invalid length')

    def itransform(self, p: np.ndarray):
        """Frenet frame -> global.  Like the reference, the SE(2) branch returns None (gn_transform_old.py:82-89)."""
        if p.shape[0] == 3:
            return None
        if p.shape[0] == 2:
            return np.matmul(self.r_mat, p) + self.t_vect
        raise ValueError('input array has invalid length')
