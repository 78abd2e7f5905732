"""Full-size parity sweep of BASELINE config 5: ALL 4096 scenarios through `BatchPlanner.plan` on the GPU against the NumPy
oracle in a process pool ("all scenarios vs the vectorised oracle").

Per scenario: n_valid, n_survivors (exact), argmin_ctot / argmin_masked (exact; a different index whose cost is within 1e-6
relative counts as a tie - the north star's exception - and is listed), min_ctot / min_masked_ctot (1e-9 relative).

    python scripts/parity_sweep.py [--field spread|literal] [--scenarios 4096] [--tile 512] [--out gpurun_out/parity_sweep_spread]

Writes <out>.json (summary + every mismatch) and <out>.md (the table committed under profiles/).
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def oracle_one(args):
    p0, s0, t0, dd, fs, fd = args
    from otg_b200 import workloads as w
    from oracle import frenet_oracle as fo
    prm = fo.OracleParams.defaults("v1").with_(**w.DENSE_PARAMS)
    sp = fo.SplinePath(w.SPLINE_WX, w.SPLINE_WY)
    ox, oy = fo.frenet_to_cartesian(sp, fs, fd)
    L = fo.generate(prm, p0, s0, t0, dd=dd)
    x, y = fo.lattice_to_cartesian(sp, L)
    ok, _ = fo.check_collisions(x, y, L.nsteps, np.stack([ox, oy], 1))
    a, m = fo.first_argmin(L.ctot, L.keep), fo.first_argmin(L.ctot, ok & L.keep)
    # runner-up costs, to classify an index difference as a tie
    return dict(n_valid=int(L.keep.sum()), n_survivors=int((ok & L.keep).sum()), argmin=int(a), argmin_masked=int(m),
                min_ctot=float(L.ctot[a]), min_masked=float(L.ctot[m]) if m >= 0 else float("inf"),
                cand_ts=int(L.nsteps.sum()), ctot=L.ctot.astype(np.float64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--field", default="spread")
    ap.add_argument("--scenarios", type=int, default=4096)
    ap.add_argument("--tile", type=int, default=512)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    out = a.out or os.path.join(ROOT, "gpurun_out", f"parity_sweep_{a.field}")
    import torch
    import otg_b200  # noqa: F401
    from otg_b200 import workloads as w
    from otg_b200.batch import BatchPlanner
    from otg_b200.planner import TrajectoryPlannerParams
    from otg_b200.trajectory import SplineTrajectory2D
    sc = w.config5_scenarios(field=a.field)
    n = a.scenarios
    fs, fd = w.moving_obstacle_frenet(sc["obs_s"][:n], sc["obs_d"][:n], sc["obs_speed"][:n], sc["obs_offset"][:n], sc["t0"][:n, None])
    # CPU side first in the background pool, GPU side while it runs
    cores = os.cpu_count() or 1
    jobs = [(sc["p0"][i], sc["s0"][i], float(sc["t0"][i]), float(sc["dd"][i]), fs[i], fd[i]) for i in range(n)]
    t0 = time.perf_counter()
    pool = mp.get_context("spawn").Pool(cores)
    async_res = pool.map_async(oracle_one, jobs, chunksize=4)
    torch.cuda.set_device(0)
    bp = None
    recs = []
    tg = time.perf_counter()
    for lo in range(0, n, a.tile):
        hi = min(lo + a.tile, n)
        if bp is None or bp.n_scen != hi - lo:
            bp = BatchPlanner(TrajectoryPlannerParams(**w.DENSE_PARAMS), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), hi - lo, fs.shape[1], "v1")
        r = bp.plan(sc["p0"][lo:hi], sc["s0"][lo:hi], sc["t0"][lo:hi], dd=sc["dd"][lo:hi], obs_s=fs[lo:hi], obs_d=fd[lo:hi])
        recs.append(r.records.copy())
    gpu_s = time.perf_counter() - tg
    rec = np.concatenate(recs)
    ref = async_res.get()
    pool.close()
    cpu_s = time.perf_counter() - t0
    mism, ties = [], []
    worst = 0.0
    for i, (r, g) in enumerate(zip(ref, rec)):
        bad = []
        if r["n_valid"] != int(g["n_valid"]):
            bad.append(("n_valid", r["n_valid"], int(g["n_valid"])))
        if r["n_survivors"] != int(g["n_survivors"]):
            bad.append(("n_survivors", r["n_survivors"], int(g["n_survivors"])))
        for key, rk, gk, ck in (("argmin_ctot", "argmin", "argmin_ctot", "min_ctot"), ("argmin_masked", "argmin_masked", "argmin_masked", "min_masked")):
            ri, gi = r[rk], int(g[gk])
            if ri != gi:
                if ri >= 0 and gi >= 0 and abs(r["ctot"][ri] - r["ctot"][gi]) <= 1e-6 * max(abs(r["ctot"][ri]), abs(r["ctot"][gi]), 1e-300):
                    ties.append((i, key, ri, gi, float(r["ctot"][ri]), float(r["ctot"][gi])))
                else:
                    bad.append((key, ri, gi))
        for key, rv, gv in (("min_ctot", r["min_ctot"], float(g["min_ctot"])), ("min_masked_ctot", r["min_masked"], float(g["min_masked_ctot"]))):
            if np.isfinite(rv) and np.isfinite(gv):
                rel = abs(rv - gv) / max(abs(rv), abs(gv), 1e-300)
                worst = max(worst, rel)
                if rel > 1e-9:
                    bad.append((key, rv, gv))
            elif rv != gv:
                bad.append((key, rv, gv))
        if bad:
            mism.append((i, bad))
    cts = sum(r["cand_ts"] for r in ref)
    summary = dict(field=a.field, scenarios=n, candidate_timesteps=int(cts), mismatches=len(mism), near_ties=len(ties),
                   max_rel_cost_diff=worst, scenarios_with_survivors=int((rec["n_survivors"] > 0).sum()),
                   survivor_fraction=float(rec["n_survivors"].sum()) / float(rec["n_valid"].sum()),
                   gpu_seconds_all_tiles_through_plan=gpu_s, oracle_seconds=cpu_s, oracle_processes=cores,
                   gpu=torch.cuda.get_device_name(0), mismatch_list=[(i, [list(map(str, b)) for b in bad]) for i, bad in mism[:50]],
                   tie_list=ties[:50])
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(summary, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as fh:
        fh.write(f"# Full config-5 parity sweep, obstacle field `{a.field}`\n\n"
                 f"`python scripts/parity_sweep.py --field {a.field} --scenarios {n}` on {summary['gpu']}: every scenario through `BatchPlanner.plan` "
                 f"(tiles of {a.tile}) against the NumPy oracle ({cores} processes, {cpu_s:.0f} s).\n\n"
                 "| scenarios | candidate-timesteps | n_valid / n_survivors / argmin / min cost mismatches | near ties (cost within 1e-6, other index) | max rel. cost difference | scenarios with survivors | survivor fraction |\n"
                 "|---|---|---|---|---|---|---|\n"
                 f"| {n} | {cts:,} | **{len(mism)}** | {len(ties)} | {worst:.2e} | {summary['scenarios_with_survivors']} | {summary['survivor_fraction']:.4f} |\n")
        if mism:
            fh.write("\nMismatches:\n\n" + "\n".join(f"* scenario {i}: {bad}" for i, bad in mism[:50]) + "\n")
        if ties:
            fh.write("\nNear ties (scenario, key, oracle index, GPU index, oracle costs of both):\n\n" + "\n".join(f"* {t}" for t in ties[:50]) + "\n")
    print(json.dumps({k: v for k, v in summary.items() if k not in ("mismatch_list", "tie_list")}))
    return 1 if mism else 0


if __name__ == "__main__":
    sys.exit(main())
