"""Parameter containers of the four planner variants.

Same class names, attribute names and defaults as the reference:
  TrajectoryPlannerDefaultParams / TrajectoryPlannerParams                   lib/planner/trajectory_planner.py:14-67
  TrajectoryPlannerDefaultParamsDT / TrajectoryPlannerParamsDT               lib/planner/trajectory_planner_dt.py:14-64
  TrajectoryPlannerDefaultParamsDTObstacles / ...ParamsDTObstacles           lib/planner/trajectory_planner_obstacles.py:14-65
  TrajectoryPlannerDefaultParamsOptimal / TrajectoryPlannerParamsOptimal     lib/planner/trajectory_planner_optimal.py:14-67
A params object copies the class-level defaults into instance attributes, then applies keyword
overrides; a planner copies a params object's `__dict__` into its own, so every parameter stays a
plain, mutable planner attribute that is re-read at the next replan.
"""
from __future__ import annotations

_COMMON = dict(kj=0.001, kt=0.01, kdots=1, klong=1, klat=1, delta_t=0.1, min_t=1, s_threshold=1, target_distance=1,
               num_sample=2)

_DEFAULTS = {
    "": dict(_COMMON, dt=0.5, ks=0.5, kd=0.5, desired_speed=2.5, max_road_width=4, max_t=3, d_road_width=0.5, d_d_s=1,
             low_speed_threshold=2),
    "DT": dict(_COMMON, ks=0.01, kd=1.0, desired_speed=0.5, max_road_width=0.4, max_t=2, d_road_width=0.1, d_d_s=0.5,
               low_speed_threshold=0.2),
    "DTObstacles": dict(_COMMON, ks=0.01, kd=2.0, klat=2, sampling_t=0.05, delta_t=0.5, desired_speed=0.5, max_road_width=0.5,
                        min_t=2, max_t=3, d_road_width=0.1, d_d_s=0.2, low_speed_threshold=0.2),
    "Optimal": dict(_COMMON, dt=0.8, ks=0.1, kd=0.5, desired_speed=2.0, max_road_width=2.45, min_t=0.9, max_t=5.0,
                    d_road_width=0.6, d_d_s=1, low_speed_threshold=2.0, s_threshold=1.0, num_sample=3),
}
# attributes a params *instance* carries (the reference does not copy s_threshold into instances)
_NOT_COPIED = ("s_threshold",)
_EXTRA_INSTANCE = {"Optimal": dict(S_thresh=2)}


def _make(suffix):
    defaults_cls = type(f"TrajectoryPlannerDefaultParams{suffix}", (), dict(_DEFAULTS[suffix]))
    copied = [k for k in _DEFAULTS[suffix] if k not in _NOT_COPIED]

    def __init__(self, *args, **kwargs):
        for k in copied:
            setattr(self, k, getattr(defaults_cls, k))
        self.s_target = None
        self.num_replan = 0
        for k, v in _EXTRA_INSTANCE.get(suffix, {}).items():
            setattr(self, k, v)
        self.__dict__.update(kwargs)

    def as_dict(self):
        return self.__dict__

    params_cls = type(f"TrajectoryPlannerParams{suffix}", (), {
        "__init__": __init__, "dict": as_dict, "__doc__": f"Container for TrajectoryPlanner{suffix or 'V1'} parameters"})
    return defaults_cls, params_cls


TrajectoryPlannerDefaultParams, TrajectoryPlannerParams = _make("")
TrajectoryPlannerDefaultParamsDT, TrajectoryPlannerParamsDT = _make("DT")
TrajectoryPlannerDefaultParamsDTObstacles, TrajectoryPlannerParamsDTObstacles = _make("DTObstacles")
TrajectoryPlannerDefaultParamsOptimal, TrajectoryPlannerParamsOptimal = _make("Optimal")
