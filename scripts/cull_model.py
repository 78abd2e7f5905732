"""CPU model of the collision cull (config 5 scenes): how many obstacles survive the per-chunk cull for different bounding
shapes, following the kernel's control flow (ascending obstacle groups of 32, first-blocker pruning, early exit).
Used to choose the cull geometry before spending GPU time; not part of the product or the tests."""
import sys
import numpy as np

sys.path.insert(0, ".")
import otg_b200  # noqa: F401,E402
from otg_b200 import workloads  # noqa: E402
from oracle import frenet_oracle as fo  # noqa: E402

R = 0.7
CH = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NSC = int(sys.argv[2]) if len(sys.argv) > 2 else 3


def seg_d2(px, py, ax, ay, wx, wy):
    l2 = wx * wx + wy * wy
    inv = 1.0 / l2 if l2 > 1e-30 else 0.0
    ux, uy = px - ax, py - ay
    t = np.clip((ux * wx + uy * wy) * inv, 0, 1)
    return (ux - t * wx) ** 2 + (uy - t * wy) ** 2


def main():
    sc = workloads.config5_scenarios(64)
    prm = fo.OracleParams.defaults("v1").with_(**workloads.DENSE_PARAMS)
    path = fo.SplinePath(workloads.SPLINE_WX, workloads.SPLINE_WY)
    tot = dict(chunks=0, circle=0, capsule=0, half=0, quarter=0, hits=0, groups=0)
    for b in range(NSC):
        L = fo.generate(prm, sc["p0"][b], sc["s0"][b], sc["t0"][b], dd=sc["dd"][b])
        x, y = fo.lattice_to_cartesian(path, L)
        ox, oy = fo.frenet_to_cartesian(path, sc["obs_s"][b], sc["obs_d"][b])
        for c in range(0, L.n_candidates, 7):
            n = int(L.nsteps[c])
            best = len(ox)
            for kb in range(0, n, CH):
                px, py = x[c, kb:min(n, kb + CH)], y[c, kb:min(n, kb + CH)]
                if not np.isfinite(px).all():
                    continue
                lim = best
                if lim == 0:
                    break
                tot["chunks"] += 1
                tot["groups"] += (lim + 31) // 32
                m = len(px) // 2
                r = np.sqrt(((px - px[m]) ** 2 + (py - py[m]) ** 2).max())
                dq = np.sqrt((ox[:lim] - px[m]) ** 2 + (oy[:lim] - py[m]) ** 2)
                near_circle = dq <= R + r
                wx, wy = px[-1] - px[0], py[-1] - py[0]
                dev = np.sqrt(seg_d2(px, py, px[0], py[0], wx, wy).max())
                near_caps = np.sqrt(seg_d2(ox[:lim], oy[:lim], px[0], py[0], wx, wy)) <= R + dev
                near_sub = {}
                for parts in (2, 4):
                    acc = np.zeros(lim, dtype=bool)
                    edges = np.linspace(0, len(px), parts + 1).astype(int)
                    for i in range(parts):
                        lo, hi = edges[i], max(edges[i + 1], edges[i] + 1)
                        qx, qy = px[lo:hi], py[lo:hi]
                        if len(qx) == 0:
                            continue
                        w2x, w2y = qx[-1] - qx[0], qy[-1] - qy[0]
                        dv = np.sqrt(seg_d2(qx, qy, qx[0], qy[0], w2x, w2y).max())
                        acc |= np.sqrt(seg_d2(ox[:lim], oy[:lim], qx[0], qy[0], w2x, w2y)) <= R + dv
                    near_sub[parts] = acc
                d2 = (px[:, None] - ox[None, :lim]) ** 2 + (py[:, None] - oy[None, :lim]) ** 2
                hit = (d2 <= R * R).any(axis=0)
                first = int(np.argmax(hit)) if hit.any() else -1
                upto = lim if first < 0 else first + 1      # the kernel stops at the first (lowest) hit
                tot["circle"] += int(near_circle[:upto].sum())
                tot["capsule"] += int(near_caps[:upto].sum())
                tot["half"] += int(near_sub[2][:upto].sum())
                tot["quarter"] += int(near_sub[4][:upto].sum())
                if first >= 0:
                    tot["hits"] += 1
                    best = first
    ch = tot["chunks"]
    print(f"chunk {CH}: {ch} chunks, obstacle groups/chunk {tot['groups'] / ch:.2f}, hits/chunk {tot['hits'] / ch:.3f}")
    for k in ("circle", "capsule", "half", "quarter"):
        print(f"  survivors per chunk, {k:8s}: {tot[k] / ch:.2f}")


main()
