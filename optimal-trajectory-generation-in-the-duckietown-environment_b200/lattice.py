"""Host-side lattice rules: which axes a planner variant samples (the reference's interval tuples).

These are the few lines of `initialize` / `generate_range_polynomials` that decide the SHAPE of the
lattice; they stay on the host and use the same Python `round()` (banker's rounding) and
`np.arange` as the reference, because those two functions define how many candidates exist
(SURVEY.md section 0, fact 5).
"""
from __future__ import annotations

import numpy as np

from . import _native as nat

# variant name -> (low-speed rule, T axis uses delta_t, D axis includes +W, sample period attribute)
VARIANTS = {
    "v1": dict(rule=nat.LOWSPEED_V1, t_from_delta=False, d_inclusive=False, period_attr="delta_t"),
    "dt": dict(rule=nat.LOWSPEED_OFF, t_from_delta=True, d_inclusive=False, period_attr="delta_t"),
    "dt_obstacles": dict(rule=nat.LOWSPEED_OFF, t_from_delta=True, d_inclusive=True, period_attr="sampling_t"),
    "optimal": dict(rule=nat.LOWSPEED_OPTIMAL, t_from_delta=False, d_inclusive=False, period_attr="delta_t"),
}


def di_interval(p, variant: str):
    """lib/planner/trajectory_planner.py:94; trajectory_planner_obstacles.py:92 adds one more sample."""
    W, step = p.max_road_width, p.d_road_width
    return (-W, W + step, step) if VARIANTS[variant]["d_inclusive"] else (-W, W, step)


def t_interval(p, variant: str):
    """trajectory_planner.py:95; the DT variants step the horizon by delta_t (trajectory_planner_dt.py:92)."""
    if VARIANTS[variant]["t_from_delta"]:
        return (p.min_t, p.max_t + p.delta_t, p.delta_t)
    return (p.min_t, p.max_t, p.dt)


def si_interval(p):
    """Target-offset axis of follow mode (trajectory_planner.py:119)."""
    return (round(-p.d_d_s * p.num_sample), round(+p.d_d_s * p.num_sample + p.d_d_s), p.d_d_s)


def dsi_interval(p, ds0):
    """Target-speed axis of velocity keeping (trajectory_planner.py:120)."""
    return (round(ds0 - p.d_d_s * p.num_sample), round(ds0 + p.d_d_s * p.num_sample + p.d_d_s), p.d_d_s)


def long_axis(p, ds0, follow: bool) -> np.ndarray:
    iv = si_interval(p) if follow else dsi_interval(p, ds0)
    return np.arange(iv[0], iv[1], iv[2])


def max_long_axis_len(p) -> int:
    """Upper bound of len(long_axis) over all ds0: span (2 n + 1) d_d_s, both ends rounded to integers."""
    span = 2 * p.d_d_s * p.num_sample + p.d_d_s + 1.0     # each end moves by at most 0.5 when rounded
    return int(np.ceil(span / p.d_d_s))


def sample_period(p, variant: str) -> float:
    return getattr(p, VARIANTS[variant]["period_attr"])
