"""CPU triage of a differential-fuzz finding (tests/test_gpu_parity.py::test_randomised_shapes_and_parameters_against_oracle).

Replays the fuzz generator up to a given (case, scenario) with the same seed / OTG_FUZZ_MAX_B, runs the oracle, and evaluates the
lateral polynomial a second time the way the CUDA path does (closed-form coefficients, repeated synthetic division) in NumPy.
If this CPU emulation differs from the oracle's power-form values by the same amount at the same sample, the finding is the
rounding of two evaluation orders on an ill-conditioned sample, not the kernel's row grouping.

    python scripts/fuzz_triage.py --seed 2026 --max-b 90 --case 16 --scenario 31 --field ddot_d
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frenet_oracle as fo   # noqa: E402


def draw_cases(seed, max_b, upto):
    rng = np.random.default_rng(seed)
    out = None
    for case in range(upto + 1):
        variant = ("v1", "dt", "dt_obstacles", "optimal")[case % 4]
        base = fo.OracleParams.defaults(variant)
        width = float(rng.choice([0.3, 1.0, 2.45, 4.0]))
        kw = dict(kj=float(rng.uniform(1e-4, 1e-2)), kt=float(rng.uniform(0, 0.1)), kd=float(rng.uniform(0.1, 2)),
                  ks=float(rng.uniform(0.01, 1)), kdots=float(rng.uniform(0.2, 2)), klat=float(rng.uniform(0.3, 2)),
                  klong=float(rng.uniform(0.3, 2)), max_road_width=width,
                  d_road_width=float(width * rng.choice([2.0, 0.7, 0.25, 0.11])),
                  min_t=float(rng.choice([0.05, 0.5, 1.0])), low_speed_threshold=float(rng.choice([0.2, 2.0])),
                  num_sample=int(rng.integers(1, 4)), d_d_s=float(rng.choice([1, 0.5])),
                  v_max=float(rng.choice([np.inf, 4.7])), a_max=float(rng.choice([np.inf, 2.3])))
        kw["max_t"] = kw["min_t"] + float(rng.choice([0.01, 0.6, 2.0]))
        period = float(rng.choice([0.01, 0.05, 0.13, 0.5]))
        if variant == "dt_obstacles":
            kw["sampling_t"] = period
            kw["delta_t"] = float(rng.choice([0.3, 0.5]))
        elif variant == "dt":
            kw["delta_t"] = max(period, 0.05)
        else:
            kw["delta_t"] = period
            kw["dt"] = float(rng.choice([0.25, 0.5, 0.8]))
        prm = base.with_(**kw)
        B = int(rng.integers(1, max_b + 1))
        p0 = np.stack([rng.uniform(-0.5, 0.5, B) * width, rng.uniform(-.5, .5, B), rng.uniform(-.5, .5, B)], 1)
        s0 = np.stack([rng.uniform(0, 30, B), rng.uniform(0, 4, B), rng.uniform(-.5, .5, B)], 1)
        dd = rng.uniform(-1, 1, B) * width
        t0 = rng.uniform(0, 3, B)
        n_T = len(np.arange(*prm.t_interval()))
        follow = bool(case % 3 == 0)
        targets = None
        if follow:
            targets = np.stack([s0[:, 0:1] + rng.uniform(1, 6, (B, n_T)), rng.uniform(0, 3, (B, n_T)), rng.uniform(-.3, .3, (B, n_T))], 2)
        O = int(rng.choice([0, 1, 7, 33, 40]))
        if O:
            rng.uniform(0.5, 10, (B, O)); rng.uniform(-1.5, 1.5, (B, O)); rng.uniform(size=(B, O))
        out = dict(prm=prm, B=B, p0=p0, s0=s0, dd=dd, t0=t0, targets=targets, O=O, variant=variant)
    return out


def synthetic_division(a, u):
    """Value and three derivatives the way csrc Poly::eval forms them (without the FMA's single rounding)."""
    a0, a1, a2, a3, a4, a5 = (a[..., j, None] for j in range(6))
    b4 = a5 * u + a4; b3 = b4 * u + a3; b2 = b3 * u + a2; b1 = b2 * u + a1; v0 = b1 * u + a0
    c4 = a5 * u + b4; c3 = c4 * u + b3; c2 = c3 * u + b2; v1 = c2 * u + b1
    d4 = a5 * u + c4; d3 = d4 * u + c3; v2 = 2.0 * (d3 * u + c2)
    e4 = a5 * u + d4; v3 = 6.0 * (e4 * u + d3)
    return v0, v1, v2, v3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--max-b", type=int, default=90)
    ap.add_argument("--case", type=int, default=16)
    ap.add_argument("--scenario", type=int, default=31)
    ap.add_argument("--field", default="ddot_d")
    a = ap.parse_args()
    c = draw_cases(a.seed, a.max_b, a.case)
    b, prm = a.scenario, c["prm"]
    L = fo.generate(prm, c["p0"][b], c["s0"][b], c["t0"][b], dd=c["dd"][b], target_triples=None if c["targets"] is None else c["targets"][b])
    j = ("d", "dot_d", "ddot_d", "dddot_d").index(a.field)
    with np.errstate(all="ignore"):
        arg = np.where(L.low_speed[:, None], np.abs(L.lon[:, 0, :] - c["s0"][b, 0]), L.t - c["t0"][b])[L.row_of]
        vals = synthetic_division(L.clat, arg)
        vals_ld = synthetic_division(L.clat.astype(np.longdouble), arg.astype(np.longdouble))
    ref = L.lat[:, j, :]
    valid = np.arange(ref.shape[1])[None, :] < L.nsteps[:, None]
    err = np.where(valid & L.keep[:, None], np.abs(vals[j] - ref), 0.0)
    err_ld = np.where(valid & L.keep[:, None], np.abs(vals_ld[j].astype(np.float64) - ref), 0.0)
    rowmax = np.nanmax(np.abs(np.where(valid, ref, np.nan)), axis=1)
    # magnitude of the terms the power form adds up for this derivative: what its rounding error scales with
    fac = {0: [1, 1, 1, 1, 1, 1], 1: [0, 1, 2, 3, 4, 5], 2: [0, 0, 2, 6, 12, 20], 3: [0, 0, 0, 6, 24, 60]}[j]
    with np.errstate(all="ignore"):
        terms = sum(fac[q] * np.abs(L.clat[:, q, None]) * np.abs(arg) ** max(q - j, 0) for q in range(6))
    cc, kk = np.unravel_index(np.argmax(err), err.shape)
    print(f"variant {c['variant']}, B={c['B']}, n_candidates {L.n_candidates}, horizons {np.arange(*prm.t_interval())}, period {prm.sample_period() if hasattr(prm, 'sample_period') else ''}")
    print(f"max |synthetic division - power form| for {a.field}: {err.max():.3e} at candidate {cc}, step {kk} "
          f"(value {ref[cc, kk]:.6e}, row max {rowmax[cc]:.3e}, term magnitude {terms[cc, kk]:.3e}, eps*terms {np.finfo(float).eps * terms[cc, kk]:.3e})")
    print(f"same in extended precision (close to what FMA contraction gives): {err_ld.max():.3e}")
    tol = 1e-6 * np.maximum(np.abs(ref), np.maximum(1.0, rowmax[:, None]) * 1e-3)
    print(f"samples over the test's floor (1e-6 * max(|ref|, 1e-3 * max(1, row max))): {(err > tol).sum()} (fp64 emulation), {(err_ld > tol).sum()} (extended)")


if __name__ == "__main__":
    main()
