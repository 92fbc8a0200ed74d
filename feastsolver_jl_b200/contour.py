"""Contour types and constructors -- host-side mirror of src/contour.jl.

Same names, argument order, defaults, node order and error behaviour as the
reference (`Contour`, `CircularContour`, `RectangularContour`, `CustomContour`,
the four constructors, `in_contour`, `rational_func`); the node/weight
arithmetic itself is done by the C ABI (csrc/contour.cpp) so that the Julia
shim and this module share one implementation.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib


class Contour:  # src/contour.jl:1
    def __len__(self):  # length(contour::Contour) = 1, contour.jl:24
        return 1


@dataclass
class CircularContour(Contour):  # contour.jl:3-8
    c: complex
    r: float
    nodes: np.ndarray
    weights: np.ndarray


@dataclass
class RectangularContour(Contour):  # contour.jl:10-16
    bottom_left: complex
    top_right: complex
    nodes: np.ndarray
    weights: np.ndarray

    def __post_init__(self):
        bl, tr = complex(self.bottom_left), complex(self.top_right)
        if not (bl.real < tr.real and bl.imag < tr.imag):
            raise ValueError("Invalid corners")  # contour.jl:15


@dataclass
class CustomContour(Contour):  # contour.jl:19-22
    nodes: np.ndarray
    weights: np.ndarray


def _call(fn, *args, N):
    z = np.empty(N, np.complex128)
    w = np.empty(N, np.complex128)
    rc = fn(*args, N, _lib.ptr(z), _lib.ptr(w))
    return rc, z, w


def circular_contour_trapezoidal(c, r, N=16):
    """contour.jl:26-31"""
    lib = _lib.load()
    rc, z, w = _call(lib.feast_contour_circular_trapezoidal, _lib.cplx(c), float(r), N=int(N))
    if rc:
        raise ValueError("Number of nodes must be positive")
    return CircularContour(c, r, z, w)


def circular_contour_gauss(c, r, N=16):
    """contour.jl:33-44"""
    lib = _lib.load()
    if N % 2 != 0:
        raise ValueError("Number of nodes must be multiple of 2")  # contour.jl:34
    rc, z, w = _call(lib.feast_contour_circular_gauss, _lib.cplx(c), float(r), N=int(N))
    if rc:
        raise ValueError("Number of nodes must be multiple of 2")
    return CircularContour(c, r, z, w)


def rectangular_contour_gauss(bottom_left, top_right, N=16):
    """contour.jl:47-63 (clockwise: top, right, bottom, left)"""
    lib = _lib.load()
    if N % 4 != 0:
        raise ValueError("Number of nodes must be multiple of 4")  # contour.jl:48
    rc, z, w = _call(lib.feast_contour_rectangular_gauss, _lib.cplx(bottom_left), _lib.cplx(top_right), N=int(N))
    if rc == -1:
        raise ValueError("Invalid corners")
    if rc:
        raise ValueError("Number of nodes must be multiple of 4")
    return RectangularContour(complex(bottom_left), complex(top_right), z, w)


def rectangular_contour_trapezoidal(bottom_left, top_right, N=16):
    """contour.jl:66-86"""
    lib = _lib.load()
    if N % 4 != 0:
        raise ValueError("Number of nodes must be multiple of 4")  # contour.jl:68
    rc, z, w = _call(lib.feast_contour_rectangular_trapezoidal, _lib.cplx(bottom_left), _lib.cplx(top_right), N=int(N))
    if rc == -1:
        raise ValueError("Invalid corners")
    if rc:
        raise ValueError("Number of nodes must be multiple of 4")
    return RectangularContour(complex(bottom_left), complex(top_right), z, w)


def in_contour(lam, contour, r=None):
    """contour.jl:88-100: closed disc (<=), open rectangle (<)."""
    lam = np.asarray(lam)
    if r is not None:  # in_contour(lam, c, r)
        return np.abs(lam - contour) <= r
    if isinstance(contour, CircularContour):
        return np.abs(lam - contour.c) <= contour.r
    if isinstance(contour, RectangularContour):
        bl, tr = complex(contour.bottom_left), complex(contour.top_right)
        return ((bl.real < lam.real) & (lam.real < tr.real) & (bl.imag < lam.imag) & (lam.imag < tr.imag))
    raise TypeError("no method matching in_contour for this contour type")  # contour.jl:18 TODO upstream


def rational_func(z, contour):
    """contour.jl:102-108"""
    S = 0.0 + 0.0j
    for zi, wi in zip(contour.nodes, contour.weights):
        S += wi / (zi - z)
    return S
