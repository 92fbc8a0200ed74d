"""feastsolver_jl_b200 -- B200-native contour-quadrature hot path of FEASTSolver.jl.

Python host mirror of the reference's entry points over the C ABI of
libfeast_cuda.so (include/feast_cuda.h).  Importing the package does not need a
GPU; every compute call does (there is no CPU fallback).
"""
from . import workloads  # noqa: F401
from ._lib import FeastError, load as load_library  # noqa: F401
from .contour import (CircularContour, Contour, CustomContour, RectangularContour,  # noqa: F401
                      circular_contour_gauss, circular_contour_trapezoidal, in_contour,
                      rational_func, rectangular_contour_gauss, rectangular_contour_trapezoidal)
from .feast import (FeastContext, I, beyn, block_SS, contour_estimate_eig, dual_gen_feast, feast, gen_feast,  # noqa: F401
                    ifeast, nlfeast, nlfeast_it, nlfeast_moments)
from .partition import node_owners  # noqa: F401

__all__ = [
    "feast", "gen_feast", "dual_gen_feast", "nlfeast", "ifeast", "nlfeast_it", "beyn", "block_SS", "nlfeast_moments", "contour_estimate_eig", "FeastContext", "FeastError", "I",
    "Contour", "CircularContour", "RectangularContour", "CustomContour",
    "circular_contour_trapezoidal", "circular_contour_gauss",
    "rectangular_contour_gauss", "rectangular_contour_trapezoidal",
    "in_contour", "rational_func", "node_owners", "workloads",
]
