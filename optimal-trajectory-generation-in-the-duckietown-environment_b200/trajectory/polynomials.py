"""Boundary-value polynomials used by the host side (leading-vehicle targets, tests).

Same class names / constructor signatures / method names as the reference's
`lib/trajectory/trajectory_1d.py:12-130`.  The classes obtain the free coefficients with
`np.linalg.solve` like the reference, because host-side callers compare sampled values with
`==` (`Target.get_frenet_stop_target` asserts an exactly zero end velocity).  `quintic_bvp` /
`quartic_bvp` are the closed-form solutions of the same systems, i.e. what the K1 CUDA kernel
evaluates (csrc/frenet_kernels.cu); the two agree to about 2e-14 relative
(tests/test_host_logic.py).
"""
from __future__ import annotations

import numpy as np


def quintic_bvp(p0, dp0, ddp0, p1, dp1, ddp1, T):
    """Coefficients (a0..a5) of the quintic through (p0,dp0,ddp0) at 0 and (p1,dp1,ddp1) at T."""
    a0 = p0
    a1 = dp0
    a2 = ddp0 / 2
    T2 = T * T
    T3 = T2 * T
    b0 = p1 - a0 - a1 * T - a2 * T2
    b1 = dp1 - a1 - 2 * a2 * T
    b2 = ddp1 - 2 * a2
    a3 = (10 * b0 - 4 * T * b1 + 0.5 * T2 * b2) / T3
    a4 = (-15 * b0 + 7 * T * b1 - T2 * b2) / (T3 * T)
    a5 = (6 * b0 - 3 * T * b1 + 0.5 * T2 * b2) / (T3 * T2)
    return a0, a1, a2, a3, a4, a5


def quartic_bvp(s0, ds0, dds0, ds1, dds1, T):
    """Coefficients (a0..a4) of the quartic with free end position (velocity keeping)."""
    a0 = s0
    a1 = ds0
    a2 = dds0 / 2.0
    b1 = ds1 - a1 - 2 * a2 * T
    b2 = dds1 - 2 * a2
    T2 = T * T
    a3 = (3 * b1 - T * b2) / (3 * T2)
    a4 = (T * b2 - 2 * b1) / (4 * T2 * T)
    return a0, a1, a2, a3, a4


class QuinticPolynomial:
    """1-D quintic; mirrors `lib/trajectory/trajectory_1d.py:12-71`."""

    def __init__(self, p0, dot_p0, ddot_p0, p1, dot_p1, ddot_p1, T):
        self.a0, self.a1, self.a2 = p0, dot_p0, ddot_p0 / 2
        powers = np.array([T ** n for n in range(6)])
        # rows: position, velocity, acceleration at T; columns: a3, a4, a5
        M = np.array([[powers[3], powers[4], powers[5]],
                      [3 * powers[2], 4 * powers[3], 5 * powers[4]],
                      [6 * powers[1], 12 * powers[2], 20 * powers[3]]])
        rhs = np.array([p1 - self.a0 - self.a1 * T - self.a2 * T ** 2,
                        dot_p1 - self.a1 - 2 * self.a2 * T,
                        ddot_p1 - 2 * self.a2])
        self.a3, self.a4, self.a5 = np.linalg.solve(M, rhs)

    def compute_pt(self, t):
        return (self.a0 + self.a1 * t + self.a2 * t ** 2 + self.a3 * t ** 3
                + self.a4 * t ** 4 + self.a5 * t ** 5)

    def compute_first_derivative(self, t):
        return (self.a1 + 2 * self.a2 * t + 3 * self.a3 * t ** 2
                + 4 * self.a4 * t ** 3 + 5 * self.a5 * t ** 4)

    def compute_second_derivative(self, t):
        return 2 * self.a2 + 6 * self.a3 * t + 12 * self.a4 * t ** 2 + 20 * self.a5 * t ** 3

    def compute_third_derivative(self, t):
        return 6 * self.a3 + 24 * self.a4 * t + 60 * self.a5 * t ** 2


class QuarticPolynomial:
    """1-D quartic; mirrors `lib/trajectory/trajectory_1d.py:74-130`."""

    def __init__(self, s0, dot_s0, ddot_s0, dot_s1, ddot_s1, T):
        self.a0, self.a1, self.a2 = s0, dot_s0, ddot_s0 / 2.0
        M = np.array([[3 * T ** 2, 4 * T ** 3], [6 * T, 12 * T ** 2]])   # velocity and acceleration rows; columns a3, a4
        rhs = np.array([dot_s1 - self.a1 - 2 * self.a2 * T, ddot_s1 - 2 * self.a2])
        self.a3, self.a4 = np.linalg.solve(M, rhs)

    def compute_pt(self, t):
        return self.a0 + self.a1 * t + self.a2 * t ** 2 + self.a3 * t ** 3 + self.a4 * t ** 4

    def compute_first_derivative(self, t):
        return self.a1 + 2 * self.a2 * t + 3 * self.a3 * t ** 2 + 4 * self.a4 * t ** 3

    def compute_second_derivative(self, t):
        return 2 * self.a2 + 6 * self.a3 * t + 12 * self.a4 * t ** 2

    def compute_third_derivative(self, t):
        return 6 * self.a3 + 24 * self.a4 * t
