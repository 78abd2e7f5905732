"""Reference paths (host side) and their device descriptors.

Class and method names follow the reference (`lib/trajectory/circle_trajectory.py:11-117`,
`lib/trajectory/spline_trajectory.py:9-189`, protocol `lib/trajectory/defs.py:8-27`) so the
sim drivers can use these objects unchanged.  The per-point host methods exist for the
driver's one-point-per-step uses (target position, curvature for the controller); the
bulk Frenet->Cartesian transform of a lattice runs on the GPU (K3) from the descriptor
returned by `path_descriptor()`.

Quirks of the reference that are kept on purpose:
  * the rounded rectangle's tangent is a forward chord over `dt = 0.1`
    (`circle_trajectory.py:66-72`), not an analytic derivative;
  * arc lengths below zero land in the last arc branch (`:62-64`);
  * a spline queried outside its knot range returns the first / last *knot abscissa*
    for both the value and the derivative (`spline_trajectory.py:47-50, 65-68`).
"""
from __future__ import annotations

import bisect
import math

import numpy as np

PATH_KIND_CIRCLE = 0
PATH_KIND_SPLINE = 1
PATH_KIND_PARABOLA = 2

# Layout of the double[16] block of a circle descriptor (see include/frenet_b200.h).
_CIRC_L, _CIRC_H, _CIRC_R, _CIRC_DT, _CIRC_PERIOD, _CIRC_B0 = 0, 1, 2, 3, 4, 5


class DifferentiableFunction:
    def compute_pt(self, t):
        raise NotImplementedError

    def compute_first_derivative(self, t):
        raise NotImplementedError

    def compute_second_derivative(self, t):
        raise NotImplementedError

    def compute_third_derivative(self, t):
        raise NotImplementedError

    def compute_curvature(self, t):
        return None


class Trajectory:
    def toJSON(self):
        return None


def circle_breakpoints(l: float, h: float, r: float):
    """Period and the seven branch boundaries, evaluated in the reference's operation
    order (`circle_trajectory.py:33-61`) so that borderline arc lengths pick the same branch."""
    pi = math.pi
    period = 2 * l + 2 * r * pi + 2 * h
    bounds = (
        l,
        l + 2 * r * pi / 4,
        l + 2 * r * pi / 4 + h,
        l + r * pi + h,
        2 * l + r * pi + h,
        2 * l + r * pi + h + 2 * r * pi / 4,
        2 * l + r * pi + 2 * h + 2 * r * pi / 4,
    )
    return period, bounds


class CircleTrajectory2D(Trajectory, DifferentiableFunction):
    """Rounded-rectangle track parameterised by arc length (l: straight, h: side, r: corner)."""

    def __init__(self, l: float, h: float, r: float):
        self.l = l
        self.h = h
        self.r = r
        self.dt = 0.1

    def _wrap(self, t: float) -> float:
        period, _ = circle_breakpoints(self.l, self.h, self.r)
        t = float(t)
        if t >= period:
            t = t % period
        return t

    def _segment(self, t: float) -> int:
        _, b = circle_breakpoints(self.l, self.h, self.r)
        if t >= 0 and t < b[0]:
            return 0
        for k in range(1, 7):
            if b[k - 1] <= t < b[k]:
                return k
        return 7

    def compute_pt(self, t):
        l, h, r = self.l, self.h, self.r
        _, b = circle_breakpoints(l, h, r)
        t = self._wrap(t)
        seg = self._segment(t)
        if seg == 0:
            return np.array([t, 0.0])
        if seg == 1:
            a = (t - l) / r
            return np.array([l + r * math.sin(a), r - r * math.cos(a)])
        if seg == 2:
            return np.array([l + r, t - l - 2 * r * math.pi / 4 + r])
        if seg == 3:
            a = (t - b[2]) / r
            return np.array([l + math.cos(a) * r, r + h + math.sin(a) * r])
        if seg == 4:
            return np.array([l - (t - l - r * math.pi - h), 2 * r + h])
        if seg == 5:
            a = (t - b[4]) / r
            return np.array([-math.sin(a) * r, r + h + r * math.cos(a)])
        if seg == 6:
            return np.array([-r, r + h - (t - b[5])])
        a = (t - b[6]) / r
        return np.array([-r * math.cos(a), r - r * math.sin(a)])

    def compute_first_derivative(self, t):
        return (self.compute_pt(t + self.dt) - self.compute_pt(t)) / self.dt

    def compute_second_derivative(self, t):
        h = 1e-5
        return self.compute_first_derivative(t + h / 2) - self.compute_first_derivative(t - h / 2)

    def compute_third_derivative(self, t):
        return np.array([0.0, 0.0])

    def compute_curvature(self, t):
        seg = self._segment(self._wrap(t))
        return 1.0 / self.r if seg % 2 == 1 else 0.0

    def path_descriptor(self):
        """(kind, n_knots, double[16] circle block, spline table or None)."""
        period, b = circle_breakpoints(self.l, self.h, self.r)
        blk = np.zeros(16, dtype=np.float64)
        blk[_CIRC_L], blk[_CIRC_H], blk[_CIRC_R], blk[_CIRC_DT] = self.l, self.h, self.r, self.dt
        blk[_CIRC_PERIOD] = period
        blk[_CIRC_B0:_CIRC_B0 + 7] = b
        return PATH_KIND_CIRCLE, 0, blk, None


class Spline:
    """Natural cubic spline y(x); mirrors `lib/trajectory/spline_trajectory.py:9-123`."""

    def __init__(self, x, y):
        self.x = x
        self.y = y
        self.nx = n = len(x)
        h = np.diff(x)
        self.a = [v for v in y]
        # Natural end conditions: c_0 = c_{n-1} = 0, interior rows are the usual
        # h_{i-1} c_{i-1} + 2 (h_{i-1} + h_i) c_i + h_i c_{i+1} = rhs_i.
        A = np.zeros((n, n))
        rhs = np.zeros(n)
        A[0, 0] = 1.0
        A[n - 1, n - 1] = 1.0
        for i in range(1, n - 1):
            A[i, i - 1] = h[i - 1]
            A[i, i] = 2.0 * (h[i - 1] + h[i])
            A[i, i + 1] = h[i]
            rhs[i] = 3.0 * (self.a[i + 1] - self.a[i]) / h[i] - 3.0 * (self.a[i] - self.a[i - 1]) / h[i - 1]
        self.c = np.linalg.solve(A, rhs)
        self.b, self.d = [], []
        for i in range(n - 1):
            self.d.append((self.c[i + 1] - self.c[i]) / (3.0 * h[i]))
            self.b.append((self.a[i + 1] - self.a[i]) / h[i] - h[i] * (self.c[i + 1] + 2.0 * self.c[i]) / 3.0)

    def _locate(self, t):
        return bisect.bisect(self.x, t) - 1

    def calc(self, t):
        if t < self.x[0]:
            return self.x[0]
        if t >= self.x[-1]:
            return self.x[-1]
        i = self._locate(t)
        dx = t - self.x[i]
        return self.a[i] + self.b[i] * dx + self.c[i] * dx ** 2.0 + self.d[i] * dx ** 3.0

    def calcd(self, t):
        if t < self.x[0]:
            return self.x[0]
        if t >= self.x[-1]:
            return self.x[-1]
        i = self._locate(t)
        dx = t - self.x[i]
        return self.b[i] + 2.0 * self.c[i] * dx + 3.0 * self.d[i] * dx ** 2.0

    def calcdd(self, t):
        if t < self.x[0]:
            return self.x[0]
        if t >= self.x[-1]:
            return self.x[-1]
        i = self._locate(t)
        dx = t - self.x[i]
        return 2.0 * self.c[i] + 6.0 * self.d[i] * dx


class SplineTrajectory2D(Trajectory, DifferentiableFunction):
    """2-D cubic spline over cumulative chord length; mirrors `spline_trajectory.py:126-189`."""

    def __init__(self, x, y):
        dx = np.diff(x)
        dy = np.diff(y)
        self.ds = [math.sqrt(ix ** 2 + iy ** 2) for ix, iy in zip(dx, dy)]
        s = [0]
        s.extend(np.cumsum(self.ds))
        self.s = s
        self.sx = Spline(self.s, x)
        self.sy = Spline(self.s, y)

    def compute_pt(self, s):
        return np.array([self.sx.calc(s), self.sy.calc(s)])

    def compute_first_derivative(self, s):
        return np.array([self.sx.calcd(s), self.sy.calcd(s)])

    def compute_second_derivative(self, s):
        return np.array([self.sx.calcdd(s), self.sy.calcdd(s)])

    def compute_third_derivative(self, s):
        return np.array([0.0, 0.0])

    def compute_curvature(self, s):
        dx, ddx = self.sx.calcd(s), self.sx.calcdd(s)
        dy, ddy = self.sy.calcd(s), self.sy.calcdd(s)
        return (ddy * dx - ddx * dy) / (dx ** 2 + dy ** 2)

    def calc_yaw(self, s):
        return math.atan2(self.sy.calcd(s), self.sx.calcd(s))

    def path_descriptor(self):
        """(kind, n_knots, circle block (zeros), table[n_knots][9]) with rows
        (knot, ax, bx, cx, dx, ay, by, cy, dy); b and d of the last knot are zero."""
        n = len(self.s)
        tab = np.zeros((n, 9), dtype=np.float64)
        tab[:, 0] = np.asarray(self.s, dtype=np.float64)
        for col, spl in ((1, self.sx), (5, self.sy)):
            tab[:, col] = np.asarray(spl.a, dtype=np.float64)
            tab[:n - 1, col + 1] = np.asarray(spl.b, dtype=np.float64)
            tab[:, col + 2] = np.asarray(spl.c, dtype=np.float64)
            tab[:n - 1, col + 3] = np.asarray(spl.d, dtype=np.float64)
        return PATH_KIND_SPLINE, n, np.zeros(16, dtype=np.float64), tab


class ParabolaTrajectory2D(Trajectory, DifferentiableFunction):
    """Lane-centre path of the Duckietown runs: y = a x^2 + b x + c in the robot frame, parameterised by x.  The reference
    builds it as a bare `Trajectory()` carrying lambdas (`MapperSemanticObstacles.build_trajectory`,
    lib/mapper/mapper_obstacles.py:169-186); this class has the same four callables plus a device descriptor."""
    ds = 1 / 30     # chord of compute_ortogonal_vect in lib/tests/test_mapper_planner_obstacles.py:137-142

    def __init__(self, a, b, c):
        self.a, self.b, self.c = float(a), float(b), float(c)

    @classmethod
    def fit(cls, x, y):
        """Quadratic least-squares fit through lane-centre points (mapper_obstacles.py:172)."""
        return cls(*np.polyfit(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), 2))

    def compute_pt(self, x):
        return np.array([x, self.a * x ** 2 + self.b * x + self.c])

    def compute_first_derivative(self, x):
        return np.array([1, 2 * self.a * x + self.b])

    def compute_second_derivative(self, x):
        return np.array([0, 2 * self.a])

    def compute_curvature(self, t):
        dt = 1 / 30
        x0, x1, x2 = t - dt, t, t + dt
        y0, y1, y2 = (self.compute_pt(x)[1] for x in (x0, x1, x2))
        t1 = np.arctan2(np.sin(y1 - y0), np.cos(x1 - x0))
        t2 = np.arctan2(np.sin(y2 - y1), np.cos(x2 - x1))
        return np.abs(t2 - t1) / np.linalg.norm(np.array([x1 - x0, y1 - y0]))

    def path_descriptor(self):
        block = np.zeros(16, dtype=np.float64)
        block[:4] = (self.a, self.b, self.c, self.ds)
        return PATH_KIND_PARABOLA, 0, block, None


def path_descriptor_of(trajectory):
    """Descriptor for one of our path objects or a duck-typed reference path object
    (anything exposing `l, h, r` or `s, sx, sy` the way the reference classes do).
    Other path kinds are not supported on the device and raise TypeError."""
    if hasattr(trajectory, "path_descriptor"):
        return trajectory.path_descriptor()
    if all(hasattr(trajectory, k) for k in ("l", "h", "r")):
        c = CircleTrajectory2D(trajectory.l, trajectory.h, trajectory.r)
        c.dt = getattr(trajectory, "dt", 0.1)
        return c.path_descriptor()
    if all(hasattr(trajectory, k) for k in ("s", "sx", "sy")):
        proxy = SplineTrajectory2D.__new__(SplineTrajectory2D)
        proxy.s, proxy.sx, proxy.sy = trajectory.s, trajectory.sx, trajectory.sy
        return proxy.path_descriptor()
    raise TypeError(
        f"reference path of type {type(trajectory).__name__} has no device descriptor "
        "(supported: CircleTrajectory2D, SplineTrajectory2D, ParabolaTrajectory2D)")
