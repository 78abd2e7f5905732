"""Throughput of the pipeline on the reference's DEFAULT lattice (configs 1-3: 320-384 candidates x 11..26 steps) for a large batch:
per-kernel CUDA-event times and HBM-roofline fractions, to see what the short horizons cost next to the dense config-5 lattice."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import otg_b200  # noqa: F401,E402
from otg_b200 import _native as nat  # noqa: E402
from otg_b200.batch import BatchPlanner  # noqa: E402
from otg_b200.planner import TrajectoryPlannerParams  # noqa: E402
from otg_b200.trajectory import CircleTrajectory2D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
O = 10
rng = np.random.default_rng(1)
p0 = np.stack([rng.uniform(-1, 1, B), rng.uniform(-.3, .3, B), rng.uniform(-.3, .3, B)], 1)
s0 = np.stack([rng.uniform(0, 30, B), rng.uniform(0, 4, B), rng.uniform(-.3, .3, B)], 1)
obs_s = s0[:, 0:1] + rng.uniform(2, 12, (B, O))
obs_d = rng.uniform(-3, 3, (B, O))
bp = BatchPlanner(TrajectoryPlannerParams(), CircleTrajectory2D(8, 5, 2), B, O, "v1", torch.device("cuda", 0))
bp.stage_inputs(p0=p0, s0=s0, t0=np.full(B, 0.1), obs_s=obs_s, obs_d=obs_d)
bp.engine.upload()
for _ in range(3):
    bp.launch_resident()
bp.engine.download(sync=True)
e = bp.engine
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L, P, Bf, path = C.byref(e.L), C.byref(bp.params), C.byref(e.B), C.byref(bp.path.desc)
calls = {
    "k1_coefficients": lambda: nat.lib.frenet_k1_coefficients(L, P, Bf, stream),
    "k2k3k4_fused": lambda: nat.lib.frenet_k2_fused(L, P, path, Bf, 1, stream),
    "k2k3_fused": lambda: nat.lib.frenet_k2_fused(L, P, path, Bf, 0, stream),
    "k2_evaluate": lambda: nat.lib.frenet_k2_evaluate(L, P, Bf, stream),
    "k5_argmin": lambda: nat.lib.frenet_k5_argmin(L, Bf, bp.use_mask, stream),
    "gather_winner": lambda: nat.lib.frenet_gather_winner(L, P, Bf, nat.SEL_MASKED, 1, stream),
}
cts = bp.cand_timesteps()
out = {"scenarios": B, "cand_timesteps": cts, "valid_candidates": int(e.result["n_valid"].sum()), "survivors": int(e.result["n_survivors"].sum())}
for name, fn in calls.items():
    nat.check(fn())
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nat.check(fn())
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out[name] = {"ms": float(np.median(ts))}
for name, per in (("k2k3k4_fused", 48), ("k2k3_fused", 48), ("k2_evaluate", 32)):
    out[name]["gbs_on_cand_ts_bytes"] = per * cts / (out[name]["ms"] * 1e-3) / 1e9
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    bp.launch_resident()
b.record()
torch.cuda.synchronize()
out["step_ms"] = a.elapsed_time(b) / 10
out["cand_ts_per_s"] = cts / (out["step_ms"] * 1e-3)
print(json.dumps(out))
