"""Synthetic problem generators for the configurations of BASELINE.json.

Every generator is deterministic (fixed numpy seeds or closed-form entries) so
that the CPU oracle and the GPU path consume identical inputs.  Shapes follow
SURVEY.md section 8(d):

C1  dense Hermitian 500x500 standard problem (scaled-up test/runtests.jl:16-20)
C2  3-D Laplacian (A) + Kronecker-sum mass matrix (B), 1-D blocks as in
    test/butterfly.jl:29-32, generalized Hermitian, analytic spectrum
C3  dense non-Hermitian complex (test/contour_random.jl:8-9 scaled up)
C4  quartic "butterfly" polynomial NEP (test/butterfly.jl:28-44, test/gen_butterfly.jl:43-64)
C5  adjacent spectral slices of the C2 pencil
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def rand_subspace(n, m0, seed=0):
    """Stand-in for Julia's rand(ComplexF64, n, m0): U[0,1) + i U[0,1)."""
    rng = np.random.default_rng(seed)
    X = np.empty((n, m0), dtype=np.complex128)
    X.real = rng.random((n, m0))
    X.imag = rng.random((n, m0))
    return X


# ---------------------------------------------------------------- C1
def dense_hermitian(n=500, seed=7):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    return (G + G.conj().T) / 2


# ---------------------------------------------------------------- C2 / C5
def _k1d(m):
    return sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(m, m), format="csr")


def _m1d(m):
    return sp.diags([1.0 / 6, 4.0 / 6, 1.0 / 6], [-1, 0, 1], shape=(m, m), format="csr")


def laplacian3d_pencil(m):
    """A = K(x)I(x)I + I(x)K(x)I + I(x)I(x)K (7-point), B = same with M, /3.
    n = m^3, nnz = 7n - 6m^2 each, identical sparsity pattern."""
    I = sp.identity(m, format="csr")
    K, M = _k1d(m), _m1d(m)

    def ksum(T):
        return (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsc()

    A = ksum(K)
    B = ksum(M) / 3.0
    A.sort_indices()
    B.sort_indices()
    return A, B


def laplacian3d_spectrum(m, count=None):
    """Exact eigenvalues of the pencil above: (k_i+k_j+k_l) / ((mu_i+mu_j+mu_l)/3),
    theta = k pi/(m+1), kappa = 2-2cos(theta), mu = (4+2cos(theta))/6."""
    th = np.arange(1, m + 1) * np.pi / (m + 1)
    kap = 2 - 2 * np.cos(th)
    mu = (4 + 2 * np.cos(th)) / 6
    if count is not None:
        # only low modes matter for the lowest slices; bound the index range
        q = min(m, max(8, int(4 * round(count ** (1 / 3))) + 8))
        kap, mu = kap[:q], mu[:q]
    num = kap[:, None, None] + kap[None, :, None] + kap[None, None, :]
    den = (mu[:, None, None] + mu[None, :, None] + mu[None, None, :]) / 3.0
    lam = np.sort((num / den).ravel())
    return lam if count is None else lam[:count]


def slice_interval(lam_sorted, first, last):
    """Circle (c, r) enclosing eigenvalues first..last-1 (0-based) of a sorted real
    spectrum, with edges at midpoints of the neighbouring gaps (5% margin below
    the lowest eigenvalue for the first slice)."""
    lo = lam_sorted[first] - 0.05 * abs(lam_sorted[first]) if first == 0 else 0.5 * (lam_sorted[first - 1] + lam_sorted[first])
    hi = 0.5 * (lam_sorted[last - 1] + lam_sorted[last])
    return 0.5 * (lo + hi), 0.5 * (hi - lo)


def c2_slice(m, target=36, first=0):
    """Pick a slice of ~target eigenvalues starting at index `first` whose upper
    edge falls in a well-separated gap.  Returns (c, r, n_inside)."""
    lam = laplacian3d_spectrum(m, count=first + 4 * target + 64)
    best = None
    for last in range(first + max(4, target - 8), first + target + 9):
        gap = lam[last] - lam[last - 1]
        if best is None or gap > best[0]:
            best = (gap, last)
    last = best[1]
    c, r = slice_interval(lam, first, last)
    return c, r, last - first


# ---------------------------------------------------------------- C3
def dense_nonhermitian(n, seed=1551):
    """randn(ComplexF64, n, n): real and imaginary parts N(0, 1/2)."""
    rng = np.random.default_rng(seed)
    A = np.empty((n, n), dtype=np.complex128)
    A.real = rng.standard_normal((n, n)) * np.sqrt(0.5)
    A.imag = rng.standard_normal((n, n)) * np.sqrt(0.5)
    return A


# ---------------------------------------------------------------- C4
_BF_C = np.array([[0.6, 1.3], [1.3, 0.1], [0.1, 1.2], [1.0, 1.0], [1.2, 1.0]])


def butterfly_coeffs(m=8, fmt="csc"):
    """Quartic butterfly coefficients M0..M4 with m x m one-dimensional blocks
    (m = 8 reproduces data/butterflyM0-M4.mtx; test/gen_butterfly.jl:43-64)."""
    N = sp.diags([1.0], [-1], shape=(m, m), format="csr")
    I = sp.identity(m, format="csr")
    Mh0 = (4 * I + N + N.T) / 6.0
    Mh1 = N - N.T
    Mh2 = -(2 * I - N - N.T)
    Mh3 = Mh1
    Mh4 = -Mh2
    out = []
    for i, Mh in enumerate([Mh0, Mh1, Mh2, Mh3, Mh4]):
        M = _BF_C[i, 0] * sp.kron(I, Mh) + _BF_C[i, 1] * sp.kron(Mh, I)
        M = M.asformat(fmt)
        M.sort_indices()
        out.append(M)
    return out
