"""Host-side trajectory primitives (same export names as the reference's `lib/trajectory/__init__.py:5-10`)."""
from .paths import (DifferentiableFunction, Trajectory, CircleTrajectory2D, Spline, SplineTrajectory2D,
                    ParabolaTrajectory2D, path_descriptor_of, PATH_KIND_CIRCLE, PATH_KIND_SPLINE, PATH_KIND_PARABOLA)
from .polynomials import QuarticPolynomial, QuinticPolynomial, quartic_bvp, quintic_bvp

__all__ = [
    "DifferentiableFunction", "Trajectory", "CircleTrajectory2D", "Spline", "SplineTrajectory2D", "ParabolaTrajectory2D",
    "QuarticPolynomial", "QuinticPolynomial", "quartic_bvp", "quintic_bvp", "path_descriptor_of",
    "PATH_KIND_CIRCLE", "PATH_KIND_SPLINE", "PATH_KIND_PARABOLA",
]
