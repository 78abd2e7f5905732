"""torchrun check of `plan_sharded` (scenario sharding + one NCCL all-gather): every rank must end up with the records
a single process computes for all scenarios.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_plan_sharded.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import otg_b200  # noqa: E402,F401
from otg_b200 import workloads as w  # noqa: E402
from otg_b200.batch import BatchPlanner, plan_sharded  # noqa: E402
from otg_b200.planner import TrajectoryPlannerParams  # noqa: E402
from otg_b200.trajectory import SplineTrajectory2D  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
n = 37                                            # ragged shards on purpose
sc = w.config5_scenarios()
fs, fd = w.moving_obstacle_frenet(sc["obs_s"][:n], sc["obs_d"][:n], sc["obs_speed"][:n], sc["obs_offset"][:n], sc["t0"][:n, None])
scen = dict(p0=sc["p0"][:n], s0=sc["s0"][:n], t0=sc["t0"][:n], dd=sc["dd"][:n], obs_s=fs, obs_d=fd)
params = TrajectoryPlannerParams(**w.DENSE_PARAMS)
tr = SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY)
records, tiles = plan_sharded(params, tr, scen, tile=8)
single = BatchPlanner(params, tr, n, 128).plan(**scen).records
ok = np.array_equal(records, single)
# equal shards: the on-device all-gather of BatchPlanner.gather_records
from otg_b200.batch import shard_range  # noqa: E402
m = 8 * world
lo, hi = shard_range(m, rank, world)
bp = BatchPlanner(params, tr, hi - lo, 128)
sub = {k: v[lo:hi] for k, v in scen.items()}
bp.stage_inputs(**sub)
bp.launch()
allrec = bp.gather_records(m).copy()
ok = ok and np.array_equal(allrec, single[:m])
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("plan_sharded", "OK" if int(t) == 1 else "MISMATCH", "world", world, "records", len(records),
          "feasible", int((records["status"] == 0).sum()))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(t) == 1 else 1)
