"""Target construction for following / merging / stopping (same exports as lib/optimal_frenet_frame)."""
from .target import Target
from .demo_parameters import OptimalParameters

__all__ = ["Target", "OptimalParameters"]
