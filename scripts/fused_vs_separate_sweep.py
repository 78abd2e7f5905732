"""ALL 4096 scenarios of config 5, both obstacle fields: the pipeline kernel (row-level collision table, exact-test candidates
compacted) against the separate launches (K2, K3, the standalone K4 that sweeps every candidate, K5) on the GPU - result records,
validity flags, collision masks, first blockers and costs of every candidate, bit for bit.  The separate launches are the form the
full oracle sweeps validated (profiles/r2_parity_sweep_*.md); this run ties the final pipeline kernel to them at full size in seconds.

    python scripts/fused_vs_separate_sweep.py > profiles/r2_fused_vs_separate_sweep.md
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import otg_b200  # noqa: E402,F401
from otg_b200 import _native as nat, workloads as w  # noqa: E402
from otg_b200.batch import BatchPlanner  # noqa: E402
from otg_b200.planner import TrajectoryPlannerParams  # noqa: E402
from otg_b200.trajectory import SplineTrajectory2D  # noqa: E402

tile = 256
print("# Pipeline kernel vs separate launches, all 4096 scenarios of config 5\n")
print("| field | scenarios | candidates compared | blocked candidates | record mismatches | mask mismatches | first-blocker mismatches | cost mismatches |")
print("|---|---|---|---|---|---|---|---|")
for field in ("spread", "literal"):
    sc = w.config5_scenarios(field=field)
    n = len(sc["t0"])
    fs, fd = w.moving_obstacle_frenet(sc["obs_s"], sc["obs_d"], sc["obs_speed"], sc["obs_offset"], sc["t0"][:, None])
    tr = SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY)
    prm = TrajectoryPlannerParams(**w.DENSE_PARAMS)
    a = BatchPlanner(prm, tr, tile, fs.shape[1], "v1")
    b = BatchPlanner(prm, tr, tile, fs.shape[1], "v1")
    b.stages |= nat.STAGE_NO_FUSE
    tot = dict(cand=0, blocked=0, rec=0, mask=0, first=0, cost=0)
    for lo in range(0, n, tile):
        s = slice(lo, lo + tile)
        kw = dict(p0=sc["p0"][s], s0=sc["s0"][s], t0=sc["t0"][s], dd=sc["dd"][s], obs_s=fs[s], obs_d=fd[s])
        ra, rb = a.plan(**kw), b.plan(**kw)
        ea, eb = a.engine, b.engine
        valid = (eb.cand_flags & nat.CAND_VALID) != 0
        tot["cand"] += int(valid.sum())
        tot["rec"] += int((np.asarray(ra.records) != np.asarray(rb.records)).sum())
        tot["rec"] += int((ea.cand_flags != eb.cand_flags).sum())
        tot["mask"] += int(((ea.coll_free != eb.coll_free) & valid).sum())
        tot["first"] += int(((ea.first_blocker != eb.first_blocker) & valid).sum())
        tot["cost"] += int(((ea.cost != eb.cost) & valid[None]).sum())
        tot["blocked"] += int(((eb.coll_free == 0) & valid).sum())
    print(f"| {field} | {n} | {tot['cand']} | {tot['blocked']} | {tot['rec']} | {tot['mask']} | {tot['first']} | {tot['cost']} |")
    del a, b
    torch.cuda.empty_cache()
