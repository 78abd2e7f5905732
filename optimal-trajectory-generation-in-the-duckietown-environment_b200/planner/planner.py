"""Planner interface (lib/planner/planner.py:7-13)."""
from abc import ABC, abstractmethod


class Planner(ABC):
    @abstractmethod
    def step(self, t: float, dd: float = None, dsd: float = None, s_target=None):
        """Frenet target position at time t (the reference's concrete planners leave this empty)."""
