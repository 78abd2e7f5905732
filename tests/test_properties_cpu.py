"""Property tests (hypothesis) of the host-side rules that fix the SHAPE of the lattice and of the oracle's invariants:
the parity hazards of SURVEY.md section 8a (np.arange lengths, banker's rounding, overshoot sample, mirror ties,
first-wins argmin)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import otg_b200  # noqa: F401
from otg_b200 import lattice as lat
from otg_b200.engine import LatticeGeometry
from otg_b200.planner import TrajectoryPlannerParams
from oracle import frenet_oracle as fo

finite = dict(allow_nan=False, allow_infinity=False)


@settings(max_examples=200, deadline=None)
@given(ds0=st.floats(-3, 8, **finite), q=st.sampled_from([1, 0.5, 0.2, 0.25, 2]), n=st.integers(1, 4))
def test_long_axis_is_numpy_arange_of_python_rounds(ds0, q, n):
    p = TrajectoryPlannerParams(d_d_s=q, num_sample=n)
    a = lat.long_axis(p, ds0, False)
    assert np.array_equal(a, np.arange(round(ds0 - q * n), round(ds0 + q * n + q), q))
    assert len(a) <= lat.max_long_axis_len(p)      # can be EMPTY (e.g. ds0 = 0, d_d_s = 0.2, num_sample = 1): the reference then has no candidates
    # the batched planner's vectorised form and the device's advance_state use rint() and (lo + q) - lo
    lo, hi = np.rint(ds0 - q * n), np.rint(ds0 + q * n + q)
    assert lo == round(ds0 - q * n) and hi == round(ds0 + q * n + q)
    nv = int(np.ceil((hi - lo) / q))
    assert nv == len(a) and np.array_equal(lo + np.arange(nv) * ((lo + q) - lo), a)
    f = lat.long_axis(p, ds0, True)
    assert np.array_equal(f, np.arange(round(-q * n), round(q * n + q), q))


@settings(max_examples=100, deadline=None)
@given(min_t=st.floats(0.3, 2.0, **finite), span=st.floats(0.2, 3.0, **finite), dt=st.sampled_from([0.1, 0.25, 0.5, 0.8]),
       period=st.sampled_from([0.02, 0.05, 0.07, 0.1, 0.13]))
def test_geometry_step_counts_padding_and_power_table(min_t, span, dt, period):
    g = LatticeGeometry.from_intervals((-1.0, 1.0, 0.5), (min_t, min_t + span, dt), period)
    assert g.n_t == len(np.arange(min_t, min_t + span, dt))
    for T, n, npad in zip(g.T, g.n_steps, g.n_pad):
        t = np.arange(0, T + period, period)
        assert n == len(t) and t[-1] >= T - 1e-9 and npad % 4 == 0 and 0 <= npad - n < 4
        assert np.array_equal(t, np.arange(n) * np.float64(period))          # sample k is exactly k * period
    assert g.t_offset[0] == 0 and np.array_equal(np.diff(g.t_offset), g.n_pad) and g.n_pad_max == g.n_pad.max()
    k = int(g.n_pad_max) - 1
    tk = float(np.float64(k) * np.float64(period))
    assert g.t_pow[k].tolist() == [tk ** 2, tk ** 3, tk ** 4, tk ** 5]


@settings(max_examples=40, deadline=None)
@given(ds0=st.floats(2.1, 5.0, **finite), s0=st.floats(0, 20, **finite), dds0=st.floats(-0.5, 0.5, **finite))
def test_mirror_candidates_tie_exactly_and_first_wins(ds0, s0, dds0):
    """With a centred start (p0 = 0, dd = 0) the +d / -d candidates have bit-identical costs; argmin takes the lower index."""
    prm = fo.OracleParams.defaults("v1")
    L = fo.generate(prm, (0.0, 0.0, 0.0), (s0, ds0, dds0), 0.1)
    nD = len(L.D)
    ctot = L.ctot.reshape(-1, nD)
    for j in range(1, nD // 2):
        assert np.array_equal(ctot[:, j], ctot[:, nD - j])
    w = fo.first_argmin(L.ctot)
    assert L.ctot[w] == L.ctot.min() and not (L.ctot[:w] == L.ctot[w]).any()
    mask = np.zeros(L.n_candidates, dtype=bool)
    assert fo.first_argmin(L.ctot, mask) == -1
    mask[[5, w]] = True
    assert fo.first_argmin(L.ctot, mask) in (5, w)


@settings(max_examples=60, deadline=None)
@given(t=st.floats(-1, 12, **finite), t0=st.floats(0, 5, **finite), n=st.integers(3, 40), period=st.sampled_from([0.05, 0.1, 0.5]))
def test_optimal_at_time_lookup_rule(t, t0, n, period):
    """trajectory_planner.py:183-195: clamp below / above, else Python round() of (time - t[0]) / period."""
    tt = np.arange(n) * period + t0
    x = np.arange(n, dtype=np.float64)
    got = fo.optimal_at_time(t, tt, x, x, x, period)[0]
    if t <= tt[0]:
        assert got == 0
    elif t >= tt[-1]:
        assert got == n - 1
    else:
        k = round((t - tt[0]) / period)
        assert got == x[k] and k == int(np.rint((t - tt[0]) / period))       # the device uses rint(): same half-to-even rule
