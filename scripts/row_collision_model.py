"""CPU model of the row-level collision stage of the fused pipeline kernel (csrc/frenet_kernels.cu, `row_collision_*`).

The n_D candidates of a lattice row share the path frame (p_k, n_k) and their lateral polynomials are affine in the lateral target:
d_i(k) = A_k + B_k * D_i.  With b = t_k . (q - p_k), a = n_k . (q - p_k) the squared distance of candidate i to obstacle q at step k
is b^2 + (a - A_k - B_k D_i)^2, so each (step, obstacle) pair blocks a contiguous RANGE of candidate indices.  The kernel computes
those ranges in fp32 with explicit error margins: inside the inner range a hit is certain, outside the outer range a miss is
certain, and only candidates in between are re-tested with the reference's fp64 expression.  This script replays that logic in
NumPy (float32 where the kernel uses fp32) on config-5 scenes and checks both inclusions against the brute-force fp64 test.

    python scripts/row_collision_model.py [--scenarios 4] [--field default|literal]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frenet_oracle as fo   # noqa: E402
import otg_b200  # noqa: E402,F401
from otg_b200 import workloads as w   # noqa: E402

f32 = np.float32
U = f32(2.0 ** -24)


def classify_row(px, py, nx, ny, A, B, D, M, obs, R2, N):
    """Returns (certain [nD][O] bool, maybe [nD][O] bool, relevant pairs) for one row; arrays over k are [N].
    Mirrors row_table_chunk: per chunk of 32 steps and obstacle ONE outer interval (hull of the lanes' outer ranges) and ONE
    inner interval (the union of the lanes' inner ranges when they form a connected run, else none)."""
    nD, O = len(D), len(obs)
    D0, dD = D[0], (D[-1] - D[0]) / max(nD - 1, 1)
    Dmax = max(abs(D[0]), abs(D[-1]))
    dx = (obs[None, :, 0] - px[:, None]).astype(f32)          # fp64 difference, rounded once
    dy = (obs[None, :, 1] - py[:, None]).astype(f32)
    nxf, nyf = nx.astype(f32)[:, None], ny.astype(f32)[:, None]
    b = dx * nyf - dy * nxf
    a = dx * nxf + dy * nyf
    Af, Bf = A.astype(f32)[:, None], B.astype(f32)[:, None]
    with np.errstate(all="ignore"):
        S = np.abs(dx) + np.abs(dy) + np.abs(Af) + np.abs(Bf) * f32(Dmax)
        delta = f32(8) * U * S + (f32(2e-12) * f32(min(M.max(), 1e15)) + f32(1e-9))
        R = f32(np.sqrt(R2))
        tau = f32(4.5) * (R + delta) * delta + f32(16) * U * f32(R2)
        g_out = f32(R2) + tau - b * b
        g_in = f32(R2) - tau - b * b
    rel = g_out >= 0
    c = Af - a                                            # d_i - a = c + B D_i
    certain = np.zeros((nD, O), bool)
    maybe = np.zeros((nD, O), bool)
    n_rel = int(rel.sum())
    idx = np.arange(nD)
    nd1 = f32(nD - 1)
    for ch in range((N + 31) // 32):
        ks = np.arange(ch * 32, min(ch * 32 + 32, N))
        for o in np.nonzero(rel[ks].any(axis=0))[0]:
            lo_m, hi_m, lo_c, hi_c = [], [], [], []
            for k in ks:
                lm, hm, lc, hc = 255, -1, 255, -1
                if rel[k, o]:
                    with np.errstate(all="ignore"):
                        r_out = np.sqrt(max(g_out[k, o], f32(0))) * f32(1.000001)
                        r_in = np.sqrt(g_in[k, o]) * f32(0.999999) if g_in[k, o] > 0 else f32(-1)
                        Bk, ck = Bf[k, 0], c[k, o]
                        if nD == 1 or not (Bk >= f32(1e-6)):
                            spread = abs(Bk) * f32(Dmax) * f32(1.000001)
                            if max(abs(ck) - spread, f32(0)) <= r_out:
                                lm, hm = 0, nD - 1
                            if abs(ck) + spread <= r_in:
                                lc, hc = 0, nD - 1
                        else:
                            inv = f32(1) / Bk * f32(1.000001)      # stands for rcp.approx
                            invd = f32(1) / f32(dD)
                            eps = f32(16) * U * invd * ((r_out + abs(ck)) * inv + f32(Dmax) * f32(1.000001)) + f32(1e-4)
                            x_lo = ((-r_out - ck) * inv - f32(D0)) * invd
                            x_hi = ((r_out - ck) * inv - f32(D0)) * invd
                            lm = int(np.ceil(min(max(x_lo - eps, f32(0)), nd1 + 1)))
                            hm = int(np.floor(max(min(x_hi + eps, nd1), f32(-1))))
                            if r_in >= 0:
                                y_lo = ((-r_in - ck) * inv - f32(D0)) * invd
                                y_hi = ((r_in - ck) * inv - f32(D0)) * invd
                                lc = int(np.ceil(min(max(y_lo + eps, f32(0)), nd1 + 1)))
                                hc = int(np.floor(max(min(y_hi - eps, nd1), f32(-1))))
                    if lm > hm:
                        lm, hm = 255, -1
                    if lc > hc:
                        lc, hc = 255, -1
                lo_m.append(lm); hi_m.append(hm); lo_c.append(lc); hi_c.append(hc)
            LO, HI = min(lo_m), max(hi_m)
            maybe[:, o] |= (idx >= LO) & (idx <= HI)
            ne = [l <= h for l, h in zip(lo_c, hi_c)]
            if any(ne):
                where = np.nonzero(ne)[0]
                contiguous = where[-1] - where[0] + 1 == len(where)
                ok = contiguous and all(not (lo_c[q + 1] > hi_c[q] + 1 or lo_c[q] > hi_c[q + 1] + 1) for q in where[:-1])
                if ok:
                    certain[:, o] |= (idx >= min(lo_c)) & (idx <= max(hi_c))
    return certain, maybe, n_rel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=3)
    ap.add_argument("--field", default="default")
    ap.add_argument("--rows", type=int, default=0, help="rows per scenario to model (0 = all)")
    args = ap.parse_args()
    prm = fo.OracleParams.defaults("v1").with_(**w.DENSE_PARAMS)
    sp = fo.SplinePath(w.SPLINE_WX, w.SPLINE_WY)
    sc = w.config5_scenarios(field=args.field) if args.field != "default" else w.config5_scenarios()
    R2 = fo.ROBOT_RADIUS ** 2
    tot = dict(rows=0, wild=0, rel=0, pairs=0, uncertain=0, cand=0, cand_with_maybe=0, viol=0, blocked=0)
    for bb in range(args.scenarios):
        fs, fd = w.moving_obstacle_frenet(sc["obs_s"][bb], sc["obs_d"][bb], sc["obs_speed"][bb], sc["obs_offset"][bb], sc["t0"][bb])
        ox, oy = fo.frenet_to_cartesian(sp, fs, fd)
        obs = np.stack([ox, oy], 1)
        L = fo.generate(prm, sc["p0"][bb], sc["s0"][bb], sc["t0"][bb], dd=sc["dd"][bb])
        x, y = fo.lattice_to_cartesian(sp, L)
        nD = len(L.D)
        Rn = len(L.V) * len(L.T)
        rows = range(Rn) if not args.rows else np.linspace(0, Rn - 1, args.rows).astype(int)
        for r in rows:
            N = int(L.nsteps[r * nD])
            s = L.lon[r, 0, :N]
            pxx, pyy = sp.pt(s)
            gx, gy = sp.tangent(s)
            th = np.arctan2(gy, gx)
            nx, ny = -np.sin(th), np.cos(th)
            arg = np.abs(s - sc["s0"][bb, 0]) if L.low_speed[r] else (L.t[r, :N] - sc["t0"][bb])
            c0, c1 = L.clat[r * nD], L.clat[(r + 1) * nD - 1]
            d_first = fo.poly_eval(c0[None], arg)[0][0]
            d_last = fo.poly_eval(c1[None], arg)[0][0]
            Bk = (d_last - d_first) / (L.D[-1] - L.D[0])
            Ak = d_first - Bk * L.D[0]
            with np.errstate(all="ignore"):
                M = np.maximum(sum(np.abs(c0[j]) * np.abs(arg) ** j for j in range(6)), sum(np.abs(c1[j]) * np.abs(arg) ** j for j in range(6)))
            tot["rows"] += 1
            Dmax = max(abs(L.D[0]), abs(L.D[-1]))
            if not (np.all(np.isfinite(Ak)) and np.all(np.isfinite(Bk)) and (np.abs(Ak) + np.abs(Bk) * Dmax).max() <= 1e12 and M.max() <= 1e15):
                tot["wild"] += 1
                continue
            certain, maybe, n_rel = classify_row(pxx, pyy, nx, ny, Ak, Bk, L.D, M, obs, R2, N)
            xr, yr = x[r * nD:(r + 1) * nD, :N], y[r * nD:(r + 1) * nD, :N]
            hit = ((xr[:, :, None] - obs[None, None, :, 0]) ** 2 + (yr[:, :, None] - obs[None, None, :, 1]) ** 2 <= R2).any(axis=1)   # [nD][O]
            tot["viol"] += int((certain & ~hit).sum() + (hit & ~(maybe | certain)).sum())
            tot["rel"] += n_rel
            tot["pairs"] += N * len(obs)
            first_certain = np.where(certain.any(1), certain.argmax(1), 10 ** 6)
            unc = maybe & ~certain & (np.arange(len(obs))[None, :] < first_certain[:, None])
            tot["uncertain"] += int(unc.sum())
            tot["cand"] += nD
            tot["cand_with_maybe"] += int(unc.any(1).sum())
            tot["blocked"] += int(hit.any(1).sum())
    print(tot)
    print(f"relevant (step, obstacle) pairs per row: {tot['rel'] / max(tot['rows'] - tot['wild'], 1):.1f} of {tot['pairs'] / max(tot['rows'] - tot['wild'], 1):.0f}; "
          f"candidates needing an exact re-test: {tot['cand_with_maybe']} of {tot['cand']} ({100 * tot['cand_with_maybe'] / max(tot['cand'], 1):.2f} %); "
          f"wild rows (fallback): {tot['wild']} of {tot['rows']}; violations: {tot['viol']}")


if __name__ == "__main__":
    main()
