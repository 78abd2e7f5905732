"""A few launches of the fp32-fast-path kernels on the bench's 512-scenario shard (for an ncu capture)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import torch
import bench
import otg_b200  # noqa: F401
from otg_b200 import _native as nat, workloads as w
from otg_b200.batch import BatchPlanner
from otg_b200.trajectory import SplineTrajectory2D
dev = torch.device("cuda", 0)
wl = bench.workload(0, 512)
bp = BatchPlanner(bench.dense_params(), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), 512, bench.N_OBS, "v1", dev, lat_f32=True)
bp.stage_inputs(**wl)
bp.engine.upload()
bp.launch_resident()
e = bp.engine
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L, P, Bf, path = C.byref(e.L), C.byref(bp.params), C.byref(e.B), C.byref(bp.path.desc)
for _ in range(3):
    nat.check(nat.lib.frenet_k2_fused(L, P, path, Bf, 0, stream))      # K2 + K3, floats stored
for _ in range(3):
    nat.check(nat.lib.frenet_k2_fused(L, P, path, Bf, 1, stream))      # + K4
torch.cuda.synchronize()
print("ok")
