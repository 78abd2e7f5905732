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


_ROW_OF = {name: i for i, name in enumerate(LIST_FIELDS)}


class WinnerFrenet(Frenet):
    """A Frenet whose list fields are made on first use from a private copy of the device's winner block.

    A replan hands back ONE such object (the pick); the next replan needs six numbers of it (`optimal_at_time`), which `_value`
    reads straight from the block.  Turning all eleven rows of a dense-lattice winner into Python lists costs ~10 us - a third of a
    single replan's host time - so that is left to whoever actually looks at a field; from then on the field IS a plain list
    (append / extend / index like the reference's), and a list that exists always wins over the block."""

    def __init__(self, rows, n_rows_valid, costs, index):
        # (no Frenet.__init__: the list slots stay unset until they are asked for)
        d = self.__dict__
        d["_rows"] = rows
        d["_item"] = rows.item                # (row, index) -> Python float
        d["_n_valid"] = n_rows_valid          # 9 without Cartesian samples, 11 with
        d["_touched"] = False                 # True once any list field exists as a list (made on demand or assigned by the caller)
        put = object.__setattr__
        put(self, "cd", costs[0]); put(self, "cv", costs[1]); put(self, "ct", costs[2]); put(self, "ctot", costs[3])
        put(self, "index", index)

    def __setattr__(self, name, value):
        if name in _ROW_OF:
            self.__dict__["_touched"] = True
        object.__setattr__(self, name, value)

    def __getattr__(self, name):              # only reached when normal lookup fails, i.e. for a list slot that is still unset
        i = _ROW_OF.get(name)
        if i is None:
            raise AttributeError(name)
        rows = self.__dict__.get("_rows")
        if rows is None:
            raise AttributeError(name)
        if i < self._n_valid:
            val = rows[i].tolist()
        else:
            # x, y of a pick that was gathered BEFORE the transform ran: frenet_to_glob leaves a fetcher behind (the reference
            # appends x, y to the very objects the planner's opt_path_* attributes point at)
            src = self.__dict__.get("_xy_source")
            val = src(name) if src is not None else []
        setattr(self, name, val)
        return val

    def _value(self, name, index):
        """field[index] as a Python float without materialising the list (a list that already exists is used instead)."""
        d = self.__dict__
        if d["_touched"]:
            try:
                return object.__getattribute__(self, name)[index]
            except AttributeError:
                pass
        return d["_item"](_ROW_OF[name], index)

    def state_at(self, time, period):
        """((d, d', d''), (s, s', s'')) at `time` - both `optimal_at_time` look-ups of a replan in one go (trajectory_planner.py:183-195,
        207-208), straight from the winner block; None if a list field has been materialised (the caller then takes the list path)."""
        d = self.__dict__
        if d["_touched"]:
            return None
        item = d["_item"]
        t_first = item(0, 0)
        if time <= t_first:
            k = 0
        elif time >= item(0, -1):
            k = -1
        else:
            k = round((time - t_first) / period)
        return (item(1, k), item(2, k), item(3, k)), (item(5, k), item(6, k), item(7, k))

    def __repr__(self):
        return f"Frenet(index={self.index}, steps={self._rows.shape[1]}, ctot={self.ctot:.6g})"
