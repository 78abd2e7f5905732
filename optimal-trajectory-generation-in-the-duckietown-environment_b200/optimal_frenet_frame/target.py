"""Leading-vehicle targets for the Optimal planner's follow mode (lib/optimal_frenet_frame/target.py:9-52).

A `Target` samples a quintic longitudinal motion every `DTs`; the follow / merge / stop variants derive the
sampled `Frenet` the planner indexes at `t_initial + T` (trajectory_planner_optimal.py:130-133).  Host-side
input producer: the planner uploads one (s, ds, dds) triple per horizon.
"""
from __future__ import annotations

import copy

import numpy as np

from ..planner.frenet import Frenet
from ..trajectory.polynomials import QuinticPolynomial


class Target:
    def __init__(self, s0=(0.0, 0.0, 0.0), s1=(1.0, 1.0, 1.0), Ts: float = 1.0):
        self.s0 = s0
        self.s1 = s1
        self.Ts = Ts       # must be the same for every Target that is combined
        self.DTs = 0.1
        self.update()

    def update(self):
        self.path = QuinticPolynomial(*self.s0, *self.s1, self.Ts)
        f = Frenet()
        f.t = [t for t in np.arange(0, self.Ts + self.DTs, self.DTs)]
        f.s = [self.path.compute_pt(t) for t in f.t]
        f.dot_s = [self.path.compute_first_derivative(t) for t in f.t]
        f.ddot_s = [self.path.compute_second_derivative(t) for t in f.t]
        f.dddot_s = [self.path.compute_third_derivative(t) for t in f.t]
        self.frenet = f

    def get_frenet(self) -> Frenet:
        return copy.deepcopy(self.frenet)

    def get_frenet_follow_target(self, D0: float = 1) -> Frenet:
        """Constant time-gap law: stay D0 + DTs * ds behind the leader (:29-36); every field is derived from
        the leader's ORIGINAL samples."""
        lead = self.frenet
        tg = self.get_frenet()
        tg.s = [s - D0 + self.DTs * v for s, v in zip(lead.s, lead.dot_s)]
        tg.dot_s = [v - self.DTs * a for v, a in zip(lead.dot_s, lead.ddot_s)]
        if not all(lead.ddot_s[0] == a for a in lead.ddot_s):
            tg.ddot_s = [a - self.DTs * j for a, j in zip(lead.ddot_s, lead.dddot_s)]
        return tg

    def get_frenet_merge_target(self, t_b=None) -> Frenet:
        assert t_b is not None, 'please give a Target'
        assert self.DTs == t_b.DTs, 'sampling times must be equal'
        assert self.Ts == t_b.Ts, 'time intervals must be equal'
        tg = self.get_frenet()
        tg.s = [0.5 * (a + b) for a, b in zip(self.frenet.s, t_b.frenet.s)]
        return tg

    def get_frenet_stop_target(self) -> Frenet:
        tg = self.get_frenet()
        assert tg.dot_s[-1] == 0, 'stopping velocity must be equal zero'
        assert tg.ddot_s[-1] == 0, 'stopping acceleration must be equal zero'
        return tg
