"""Headless versions of the reference's planner experiments (`tester.py -t planner_full | planner_obstacles |
planner_moving_obstacles`) plus the driver-level K3 / K4 / masked-K5 functions they are built from."""
from . import obstacle_planner, moving_obstacle_planner, planner_full
from .common import robot_radius, compute_ortogonal_vect, frenet_to_glob, best_path
from .planner_full import test_planner_full
from .obstacle_planner import test_planner_obstacle
from .moving_obstacle_planner import test_planner_moving_obstacle

TEST_MAP = {
    "planner_full": test_planner_full,
    "planner_obstacles": test_planner_obstacle,
    "planner_moving_obstacles": test_planner_moving_obstacle,
}
