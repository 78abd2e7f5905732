"""Demo constants of the open-loop planner exercise (lib/optimal_frenet_frame/utils.py:7-22)."""
import numpy as np


class OptimalParameters:
    p = (3, 0.3, 0)          # initial lateral state
    s = (0, 2, -0.5)         # initial longitudinal state
    s0_t = (1, 2, 0)         # leading vehicle: start / end states, duration
    s1_t = (8, 0, 0)
    Ts = 10
    t0 = 0                   # planner initialisation time
    Tn = [0, 2.5, 5]         # replanning instants
    T = np.arange(0, 5, 0.5)
    s0_tb = (3, 2, 0)        # second vehicle (merge)
    s1_tb = (10, 0, 0)
    px = (3, 0.2, 0)
    sx = (0.8, 2, -0.5)


op = OptimalParameters
