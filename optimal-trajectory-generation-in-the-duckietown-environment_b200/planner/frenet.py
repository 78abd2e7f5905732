"""`Frenet`: one candidate trajectory as the callers of the reference see it (lib/planner/frenet.py:2-19).

Plain lists and floats, exactly like the reference's container, so driver code that indexes, appends
to, zips or deep-copies the fields keeps working.  Instances are produced on demand from the device
lattice (see `LatticePaths`); nothing on the hot path builds them.
"""
from __future__ import annotations

LIST_FIELDS = ("t", "d", "dot_d", "ddot_d", "dddot_d", "s", "dot_s", "ddot_s", "dddot_s", "x", "y")
COST_FIELDS = ("cd", "cv", "ctot", "ct")


class Frenet:
    __slots__ = LIST_FIELDS + COST_FIELDS + ("index",)

    def __init__(self):
        for name in LIST_FIELDS:
            setattr(self, name, [])
        for name in COST_FIELDS:
            setattr(self, name, 0.0)
        self.index = -1     # position in the lattice's candidate list (not part of the reference's class)

    def __repr__(self):
        return f"Frenet(index={self.index}, steps={len(self.t)}, ctot={self.ctot:.6g})"
