"""Abstract planner interface.

Mirrors the single-method ABC of the reference (`lib/planner/planner.py:7-13`): a planner answers
"where should the robot be, in Frenet coordinates, at time t".  The lattice planners of this package
do their work in `replanner`, so their `step` is an inert override, as in the reference.
"""
import abc


class Planner(abc.ABC):
    @abc.abstractmethod
    def step(self, t, dd=None, dsd=None, s_target=None):
        raise NotImplementedError
