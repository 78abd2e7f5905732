"""Constants of the open-loop planner demo (values of `lib/optimal_frenet_frame/utils.py:7-22`).

`OptimalParameters` keeps the attribute names the demo flows use (`op.p`, `op.s0_t`, `op.Tn`, ...).
"""
import numpy as np

_STATE = {"p": (3, 0.3, 0), "s": (0, 2, -0.5), "px": (3, 0.2, 0), "sx": (0.8, 2, -0.5)}           # ego: lateral / longitudinal triples
_LEADERS = {"s0_t": (1, 2, 0), "s1_t": (8, 0, 0), "s0_tb": (3, 2, 0), "s1_tb": (10, 0, 0)}       # two leading vehicles: start / end
_TIMING = {"Ts": 10, "t0": 0, "Tn": [0, 2.5, 5], "T": np.arange(0, 5, 0.5)}                     # target duration, init time, replans

OptimalParameters = type("OptimalParameters", (), {**_STATE, **_LEADERS, **_TIMING})
op = OptimalParameters
