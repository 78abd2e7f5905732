"""Unicycle kinematics with the reference's closed-form pose increment (lib/platform/unicycle.py:6-84)."""
from __future__ import annotations

import numpy as np


def _sinc_series(a):      # sin(a) / a, 7th-order Taylor (:24-25)
    return 1 - 1 / 6 * a ** 2 + 1 / 120 * a ** 4 - 1 / 5040 * a ** 6


def _cosc_series(a):      # (1 - cos(a)) / a (:27-28)
    return 1 / 2 * a - 1 / 24 * a ** 3 + 1 / 720 * a ** 5 - 1 / 40320 * a ** 7


def _pose_matrix(v):
    c, s = np.cos(v[2]), np.sin(v[2])
    return np.array([[c, -s, v[0]], [s, c, v[1]], [0, 0, 1]])


class Unicycle:
    def __init__(self):
        self.p = np.array([0.0, 0.0, 0.0])

    def set_initial_pose(self, p0: np.ndarray):
        # the reference keeps `p0` as an alias of `p` (both names point at the same array, :12-14), so the
        # "initial pose" used by the integrator is always the current pose
        self.p = p0
        self.p0 = p0

    def step(self, u: np.ndarray, dt: float = 0.1):
        """Apply (v, omega) for dt; returns (pose, velocity)."""
        assert u.shape == (2,)
        d_rho, d_theta = dt * u[0], dt * u[1]
        inc = np.array([d_rho * _sinc_series(d_theta), d_rho * _cosc_series(d_theta), d_theta])
        A = _pose_matrix(self.p0) @ _pose_matrix(inc)
        new = np.array([A[0, 2], A[1, 2], np.arctan2(A[1, 0], A[0, 0])])
        dp = new - self.p
        theta = np.arctan2(np.sin(dp[2:]), np.cos(dp[2:]))
        vel = np.concatenate([dp[:2], theta], axis=0) / dt
        self.p += dt * vel
        return self.p, vel

    def pose(self):
        return self.p
