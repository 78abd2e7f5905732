"""Planner package: same exports as the reference's `lib/planner/__init__.py:5-10`."""
from .base import Planner
from .frenet import Frenet
from .params import (TrajectoryPlannerDefaultParams, TrajectoryPlannerParams,
                     TrajectoryPlannerDefaultParamsDT, TrajectoryPlannerParamsDT,
                     TrajectoryPlannerDefaultParamsDTObstacles, TrajectoryPlannerParamsDTObstacles,
                     TrajectoryPlannerDefaultParamsOptimal, TrajectoryPlannerParamsOptimal)
from .lattice_planner import (TrajectoryPlannerV1, TrajectoryPlannerV1DT, TrajectoryPlannerV1DTObstacles,
                              TrajectoryPlannerOptimal, LatticePaths)

__all__ = [
    "Planner", "Frenet", "LatticePaths",
    "TrajectoryPlannerV1", "TrajectoryPlannerParams", "TrajectoryPlannerDefaultParams",
    "TrajectoryPlannerV1DT", "TrajectoryPlannerParamsDT", "TrajectoryPlannerDefaultParamsDT",
    "TrajectoryPlannerV1DTObstacles", "TrajectoryPlannerParamsDTObstacles", "TrajectoryPlannerDefaultParamsDTObstacles",
    "TrajectoryPlannerOptimal", "TrajectoryPlannerParamsOptimal", "TrajectoryPlannerDefaultParamsOptimal",
]
