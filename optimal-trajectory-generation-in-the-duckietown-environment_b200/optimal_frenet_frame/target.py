"""Leading-vehicle targets for the Optimal planner's follow mode (behaviour of `lib/optimal_frenet_frame/target.py:9-52`).

A `Target` is a quintic longitudinal motion of a leading vehicle sampled every `DTs`; the planner indexes the
sampled `Frenet` at `t_initial + T` (`lib/planner/trajectory_planner_optimal.py:130-133`).  Three derived
targets exist: follow (constant time gap), merge (midpoint of two leaders) and stop.  Everything here is a
host-side input producer: per replan the planner uploads one (s, ds, dds) triple per horizon.

The samples are produced array-wise; each element goes through the same floating-point operations as the
reference's per-sample list comprehensions, so the values are bit-identical (tests/test_gpu_planner.py).
"""
from __future__ import annotations

import numpy as np

from ..planner.frenet import Frenet
from ..trajectory.polynomials import QuinticPolynomial

_SAMPLED = ("t", "s", "dot_s", "ddot_s", "dddot_s")


def _frenet_from(arrays: dict) -> Frenet:
    f = Frenet()
    for name in _SAMPLED:
        setattr(f, name, [v for v in arrays[name]])
    return f


class Target:
    DTs = 0.1      # sampling period; must agree between targets that are combined

    def __init__(self, s0=(0.0, 0.0, 0.0), s1=(1.0, 1.0, 1.0), Ts: float = 1.0):
        self.s0, self.s1, self.Ts = s0, s1, Ts
        self.update()

    def update(self):
        """(Re)sample the leader's motion from its boundary states."""
        q = self.path = QuinticPolynomial(*self.s0, *self.s1, self.Ts)
        t = np.arange(0, self.Ts + self.DTs, self.DTs)
        evaluators = (q.compute_pt, q.compute_first_derivative, q.compute_second_derivative, q.compute_third_derivative)
        self._arr = {"t": t, **{name: np.array([fn(ti) for ti in t]) for name, fn in zip(_SAMPLED[1:], evaluators)}}
        self.frenet = _frenet_from(self._arr)

    def get_frenet(self) -> Frenet:
        return _frenet_from(self._arr)

    def get_frenet_follow_target(self, D0: float = 1) -> Frenet:
        """Constant time-gap law: D0 + DTs * (leader speed) behind the leader; the speed (and, for a leader with
        varying acceleration, the acceleration) is shifted by one sampling period of the next derivative."""
        a = self._arr
        out = dict(a)
        out["s"] = a["s"] - D0 + self.DTs * a["dot_s"]
        out["dot_s"] = a["dot_s"] - self.DTs * a["ddot_s"]
        if not np.all(a["ddot_s"] == a["ddot_s"][0]):
            out["ddot_s"] = a["ddot_s"] - self.DTs * a["dddot_s"]
        return _frenet_from(out)

    def get_frenet_merge_target(self, t_b=None) -> Frenet:
        """Midpoint between this leader and `t_b` (speeds and accelerations of this leader are kept)."""
        assert t_b is not None, 'please give a Target'
        assert self.DTs == t_b.DTs, 'sampling times must be equal'
        assert self.Ts == t_b.Ts, 'time intervals must be equal'
        out = dict(self._arr)
        out["s"] = 0.5 * (self._arr["s"] + t_b._arr["s"])
        return _frenet_from(out)

    def get_frenet_stop_target(self) -> Frenet:
        """The leader's own motion, which must end at rest."""
        assert self._arr["dot_s"][-1] == 0, 'stopping velocity must be equal zero'
        assert self._arr["ddot_s"][-1] == 0, 'stopping acceleration must be equal zero'
        return self.get_frenet()
