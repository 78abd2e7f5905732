"""Host-side pieces of the closed-loop simulation that sit either side of the planner hot path.

They are per-step scalar computations (one Gauss-Newton projection, one 2x2 solve, one unicycle step per
simulation step), out of scope for the GPU path, and restated here only so
that the headless drivers (`otg_b200.drivers`) can run the reference's closed-loop experiments.
"""
from .transform import FrenetGNTransformOld
from .controller import FrenetIOLController
from .platform import Unicycle

__all__ = ["FrenetGNTransformOld", "FrenetIOLController", "Unicycle"]
