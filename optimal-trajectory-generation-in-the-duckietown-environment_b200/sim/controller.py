"""Input-output linearisation tracking controller in Frenet coordinates (lib/controller/ctrl_frenet_iol.py:11-103)."""
from __future__ import annotations

import numpy as np


class FrenetIOLController:
    def __init__(self, b: float, kp1: float, kp2: float, kd1: float, kd2: float):
        self.b = b
        self.kp1, self.kp2, self.kd1, self.kd2 = kp1, kp2, kd1, kd2
        self.kp = np.array([kp1, kp2])
        self.kd = np.array([kd1, kd2])

    def compute(self, robot_fpose: np.ndarray, error: np.ndarray, t_fvel: np.ndarray, k: float) -> np.ndarray:
        """(v, omega) from the Frenet pose (s, d, theta~), the position error, the planner's feed-forward
        velocity and the path curvature; the error along s is ignored (:97-98)."""
        b = self.b
        d, th = robot_fpose[1], robot_fpose[2]
        c, s = np.cos(th), np.sin(th)
        jac = np.array([[c * (1 + b * k * s) / (1 - k * d), -b * s],
                        [s - (k * c) / (1 - k * d), b * c]])
        p_term = error * self.kp
        p_term[0] = 0.0
        return np.matmul(np.linalg.inv(jac), (t_fvel[:2] + p_term).T)
