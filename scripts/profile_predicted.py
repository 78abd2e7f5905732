"""Three predicted-mode plans of a config-5 batch (for ncu captures of the fused kernel's predicted-collision variant)."""
import sys

import torch

sys.path.insert(0, ".")
import otg_b200  # noqa: F401,E402
from otg_b200 import workloads as w  # noqa: E402
from otg_b200.batch import BatchPlanner  # noqa: E402
from otg_b200.planner import TrajectoryPlannerParams  # noqa: E402
from otg_b200.trajectory import SplineTrajectory2D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
predict = int(sys.argv[2]) if len(sys.argv) > 2 else 150
sc = w.config5_scenarios(n)
bp = BatchPlanner(TrajectoryPlannerParams(**w.DENSE_PARAMS), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), n, 128, "v1",
                  torch.device("cuda", 0), predict_steps=predict)
kw = dict(p0=sc["p0"], s0=sc["s0"], t0=sc["t0"], dd=sc["dd"], obs_s=sc["obs_s"], obs_d=sc["obs_d"] + sc["obs_offset"])
if predict:
    kw["obs_speed"] = sc["obs_speed"]
bp.stage_inputs(**kw)
bp.engine.upload()
for _ in range(3):
    bp.launch_resident()
bp.engine.download(sync=True)
print("feasible scenarios", int((bp.engine.result["status"] == 0).sum()), "of", n)
