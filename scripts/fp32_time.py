"""CUDA-event times of the fp32-fast-path kernels (and the fp64 ones beside them) on the bench's 512-scenario shard."""
import ctypes as C
import json
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
out = {"lib": nat.LIB_PATH.split("/")[-1]}
for f32 in (True, False):
    bp = BatchPlanner(bench.dense_params(), SplineTrajectory2D(w.SPLINE_WX, w.SPLINE_WY), 512, bench.N_OBS, "v1", dev, lat_f32=f32)
    bp.stage_inputs(**wl)
    bp.engine.upload()
    bp.launch_resident()
    e = bp.engine
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L, P, Bf, path = C.byref(e.L), C.byref(bp.params), C.byref(e.B), C.byref(bp.path.desc)
    tag = "f32" if f32 else "f64"
    out[tag + "_k2k3"] = round(bench._event_ms(lambda: nat.check(nat.lib.frenet_k2_fused(L, P, path, Bf, 0, stream)), 20), 4)
    out[tag + "_k2k3k4"] = round(bench._event_ms(lambda: nat.check(nat.lib.frenet_k2_fused(L, P, path, Bf, 1, stream)), 20), 4)
    del bp, e
    torch.cuda.empty_cache()
print(json.dumps(out))
