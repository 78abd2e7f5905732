import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, _gpu
from otg_b200 import _native as nat
from otg_b200.trajectory import CircleTrajectory2D
from oracle import frenet_oracle as fo
tr=CircleTrajectory2D(8,5,2)
for nf in (False, True):
    eng=_gpu.run_batch(fo.OracleParams.defaults("v1"),(0,0,0),(1.0,1.0,0),0.1,trajectory=tr,obstacles=np.array([[[1.0,0.0]]]), no_fuse=nf)
    print("no_fuse",nf, eng.result[0])
    f=eng.padded_fields(eng.fetch_scenario(0))
    print(f["coll_free"][:20], f["first_blocker"][:20], f["x"][0,:3], f["y"][0,:3])
    print(eng.obs_xy, eng.obs_active, eng.B.n_obs)
