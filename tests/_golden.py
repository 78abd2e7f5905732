"""Helpers shared by the oracle and GPU parity tests: golden loading and array packing."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
sys.path.insert(0, ROOT)

FIELDS = ("t", "d", "dot_d", "ddot_d", "dddot_d", "s", "dot_s", "ddot_s", "dddot_s")
LAT = ("d", "dot_d", "ddot_d", "dddot_d")
LON = ("s", "dot_s", "ddot_s", "dddot_s")
RTOL = 1e-5          # north star: coefficients, costs and poses within 1e-5 relative in fp64
ORACLE_RTOL = 1e-9   # the oracle restates the same NumPy calls; it is held to a much tighter bound

_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return _cache[name]


def prefixed(z, prefix):
    """Sub-dictionary of a golden file: keys starting with `prefix`, prefix removed."""
    return {k[len(prefix):]: v for k, v in z.items() if k.startswith(prefix)}


def pack_lattice(L, t_field=True):
    """Oracle Lattice -> the golden layout: kept candidates only, [C][Nmax] arrays, NaN padding."""
    keep = L.keep
    idx = np.nonzero(keep)[0]
    rows = L.row_of[idx]
    out = {"nsteps": L.nsteps[idx].astype(np.int32)}
    for j, name in enumerate(LAT):
        out[name] = L.lat[idx, j, :]
    for j, name in enumerate(LON):
        out[name] = L.lon[rows, j, :]
    if t_field:
        out["t"] = L.t[rows, :]
    for k in ("cd", "cv", "ct", "ctot"):
        out[k] = getattr(L, k)[idx]
    return out


def assert_close(a, b, rtol, what="", atol_scale=None, per_row=False, extra_atol=None):
    """Relative comparison scaled by the magnitude of the reference array (NaN padding must match).
    per_row: the absolute floor scales with each row's own magnitude (ill-conditioned low-speed candidates
    reach |d| ~ 1e15 next to ordinary ones in the same array).
    extra_atol: array (broadcastable) added to the tolerance - the rounding floor of the sample itself, e.g.
    64 eps x the magnitude of the terms a polynomial derivative sums up (see poly_term_bound)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN padding differs"
    if a.size == 0:
        return
    if per_row and b.ndim >= 2:
        with np.errstate(all="ignore"):
            scale = np.maximum(1.0, np.nanmax(np.abs(np.where(nb, 0.0, b)), axis=-1, keepdims=True)) * (atol_scale or 1.0)
        atol_scale = None
    else:
        scale = atol_scale if atol_scale is not None else max(1.0, float(np.nanmax(np.abs(b))) if (~nb).any() else 1.0)
    err = np.abs(np.where(na, 0.0, a - b))
    tol = rtol * np.maximum(np.abs(np.where(nb, 0.0, b)), scale * 1e-3) if atol_scale is None else rtol * scale
    if extra_atol is not None:
        tol = tol + np.nan_to_num(np.asarray(extra_atol, dtype=np.float64), nan=0.0, posinf=np.inf)
    bad = err > tol
    assert not bad.any(), f"{what}: max err {err.max():.3e} at {np.argwhere(bad)[:3].tolist()} (rtol {rtol})"


_DERIV_FACTORS = ((1, 1, 1, 1, 1, 1), (0, 1, 2, 3, 4, 5), (0, 0, 2, 6, 12, 20), (0, 0, 0, 6, 24, 60))


def poly_term_bound(coef, arg, order):
    """sum_j f_j |a_j| |u|^(j - order): the magnitude of the terms the `order`-th derivative of a polynomial adds up at u.
    Any two evaluation orders (the reference's power form, Horner, synthetic division) and any two routes to the
    coefficients (np.linalg.solve, closed form) agree only to a few eps of THIS, not of the value - which matters where the
    value is a cancellation (e.g. the end acceleration 0 of a low-speed quintic with a horizon of millimetres, whose 3x3
    system has a condition number of 1e11).  coef [..., 6], arg [..., N] -> [..., N]."""
    f = _DERIV_FACTORS[order]
    with np.errstate(all="ignore"):
        u = np.abs(np.asarray(arg, dtype=np.float64))
        return sum(f[j] * np.abs(coef[..., j, None]) * u ** max(j - order, 0) for j in range(6))


def compare_packed(got, ref, rtol, what="", fields=None):
    assert np.array_equal(got["nsteps"], ref["nsteps"]), f"{what}: step counts differ"
    n = ref["nsteps"]
    nmax = ref[next(k for k in ref if k in FIELDS)].shape[1] if any(k in ref for k in FIELDS) else 0
    for name in (fields or FIELDS):
        if name not in ref or name not in got:
            continue
        g = got[name][:, :nmax] if got[name].shape[1] >= nmax else np.pad(
            got[name], ((0, 0), (0, nmax - got[name].shape[1])), constant_values=np.nan)
        # the oracle pads rows up to its own Nmax with NaN as well; re-mask with the golden's step counts
        g = np.where(np.arange(nmax)[None, :] < n[:, None], g, np.nan)
        assert_close(g, ref[name], rtol, f"{what}.{name}")
    for k in ("cd", "cv", "ct", "ctot"):
        if k in ref and k in got:
            assert_close(got[k], ref[k], rtol, f"{what}.{k}")


def near_tie(values, idx_a, idx_b, rel=1e-6):
    """True if the two candidates' costs are within `rel` relative (the north star's tie exception)."""
    a, b = values[idx_a], values[idx_b]
    return abs(a - b) <= rel * max(abs(a), abs(b), 1e-300)
