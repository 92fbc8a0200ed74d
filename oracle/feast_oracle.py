"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.

A numpy/scipy restatement of the contour-quadrature hot path of
spacedome/FEASTSolver.jl.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker / the timed CPU arm.  The product path
(``feastsolver_jl_b200``) never imports it and has no CPU fallback.

Pinning status: the reference is Julia and cannot run in this environment, so
this restatement is pinned against every assertion the reference's own test
suite makes for the path (``test/runtests.jl:16-23,33-49`` -> tests/test_oracle.py)
and against the exact answers of the reference's ``companion()`` construction on
the shipped ``data/*.mtx`` fixtures (tests/golden/).  ``nlfeast!``,
``dual_gen_feast!``, ``circular_contour_gauss`` and B != I are **parity
unpinned** by any reference assertion; for those the
oracle is anchored on analytic spectra and companion linearisation instead.

All ``file:line`` citations are into /root/reference/.
Arithmetic lives in Julia stdlib LinearAlgebra (OpenBLAS 0.3.21 / SuiteSparse
UMFPACK, Manifest.toml:176-179,286-292) which is not vendored; the equivalents
used here are scipy's LAPACK (zgetrf/zgetrs/zgeqrf/zgeev/zggev/zgesdd) and
SuperLU for sparse factorisations.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------
# contours  (src/contour.jl)
# --------------------------------------------------------------------------


@dataclass
class CircularContour:  # src/contour.jl:3-8
    c: complex
    r: float
    nodes: np.ndarray
    weights: np.ndarray


@dataclass
class RectangularContour:  # src/contour.jl:10-16
    bottom_left: complex
    top_right: complex
    nodes: np.ndarray
    weights: np.ndarray

    def __post_init__(self):
        bl, tr = complex(self.bottom_left), complex(self.top_right)
        if not (bl.real < tr.real and bl.imag < tr.imag):  # contour.jl:15
            raise ValueError("Invalid corners")


@dataclass
class CustomContour:  # src/contour.jl:19-22 (no in_contour method upstream)
    nodes: np.ndarray
    weights: np.ndarray


def circular_contour_trapezoidal(c, r, N=16):
    """src/contour.jl:26-31: theta = LinRange(pi/N, 2pi - pi/N, N)."""
    theta = np.linspace(np.pi / N, 2 * np.pi - np.pi / N, N)
    e = np.exp(1j * theta)
    return CircularContour(c, r, r * e + c, r * e / N)


def circular_contour_gauss(c, r, N=16):
    """src/contour.jl:33-44: two Gauss-Legendre half circles."""
    if N % 2 != 0:
        raise ValueError("Number of nodes must be multiple of 2")
    n = N // 2
    x, w = np.polynomial.legendre.leggauss(n)
    phi = (np.pi / 2.0) * (x + 1.0)
    nodes = np.zeros(N, complex)
    weights = np.zeros(N, complex)
    nodes[:n] = r * np.exp(1j * phi) + c
    nodes[n:] = r * np.exp(1j * (phi + np.pi)) + c
    weights[:n] = r * np.exp(1j * phi) * w / 4.0
    weights[n:] = r * np.exp(1j * (phi + np.pi)) * w / 4.0
    return CircularContour(c, r, nodes, weights)


def rectangular_contour_gauss(bottom_left, top_right, N=16):
    """src/contour.jl:47-63: clockwise top, right, bottom, left."""
    if N % 4 != 0:
        raise ValueError("Number of nodes must be multiple of 4")
    n = N // 4
    bl, tr = complex(bottom_left), complex(top_right)
    x, w = np.polynomial.legendre.leggauss(n)
    top_len, side_len = tr.real - bl.real, tr.imag - bl.imag
    nodes = np.zeros(N, complex)
    weights = np.zeros(N, complex)
    nodes[0:n] = (x + 1) * (top_len / 2) + (tr.imag * 1j + bl.real)
    nodes[n:2 * n] = (x + 1) * (1j * side_len / 2) + (bl.imag * 1j + tr.real)
    nodes[2 * n:3 * n] = (x[::-1] + 1) * (top_len / 2) + (bl.imag * 1j + bl.real)
    nodes[3 * n:4 * n] = (x[::-1] + 1) * (1j * side_len / 2) + (bl.imag * 1j + bl.real)
    weights[0:n] = w * top_len
    weights[n:2 * n] = -1j * w * side_len
    weights[2 * n:3 * n] = -w * top_len
    weights[3 * n:4 * n] = 1j * w * side_len
    return RectangularContour(bl, tr, nodes, weights / (-4.0 * np.pi * 1j))


def rectangular_contour_trapezoidal(bottom_left, top_right, N=16):
    """src/contour.jl:66-86: equispaced points per side, half weights at corners."""
    if N % 4 != 0:
        raise ValueError("Number of nodes must be multiple of 4")
    n = N // 4
    bl, tr = complex(bottom_left), complex(top_right)
    nodes = np.zeros(N, complex)
    weights = np.zeros(N, complex)
    nodes[0:n] = np.linspace(bl.real, tr.real, n + 1)[:n] + tr.imag * 1j
    nodes[n:2 * n] = np.linspace(tr.imag, bl.imag, n + 1)[:n] * 1j + tr.real
    nodes[2 * n:3 * n] = np.linspace(tr.real, bl.real, n + 1)[:n] + bl.imag * 1j
    nodes[3 * n:4 * n] = np.linspace(bl.imag, tr.imag, n + 1)[:n] * 1j + bl.real
    top_len, side_len = tr.real - bl.real, tr.imag - bl.imag
    weights[0] = 1j * side_len / (2 * n) + top_len / (2 * n)
    weights[1:n] = top_len / n
    weights[n] = top_len / (2 * n) - 1j * side_len / (2 * n)
    weights[n + 1:2 * n] = -1j * side_len / n
    weights[2 * n] = -1j * side_len / (2 * n) - top_len / (2 * n)
    weights[2 * n + 1:3 * n] = -top_len / n
    weights[3 * n] = -top_len / (2 * n) + 1j * side_len / (2 * n)
    weights[3 * n + 1:4 * n] = 1j * side_len / n
    return RectangularContour(bl, tr, nodes, weights / (-2.0 * np.pi * 1j))


def in_contour(lam, contour, r=None):
    """src/contour.jl:88-100.  Circle is CLOSED (<=), rectangle is STRICT (<)."""
    lam = np.asarray(lam)
    if r is not None:  # in_contour(lam, c, r) form, contour.jl:88-90
        return np.abs(lam - contour) <= r
    if isinstance(contour, CircularContour):
        return np.abs(lam - contour.c) <= contour.r
    if isinstance(contour, RectangularContour):
        bl, tr = complex(contour.bottom_left), complex(contour.top_right)
        return ((bl.real < lam.real) & (lam.real < tr.real)
                & (bl.imag < lam.imag) & (lam.imag < tr.imag))
    raise TypeError("in_contour is not defined for this contour type (contour.jl:18)")


def rational_func(z, contour):
    """src/contour.jl:102-108: sum_i w_i / (z_i - z)."""
    return np.sum(contour.weights / (contour.nodes - z))


# --------------------------------------------------------------------------
# helpers  (src/utils.jl)
# --------------------------------------------------------------------------

def _is_sparse(A):
    return sp.issparse(A)


def _julia_eig_sort(w, v):
    """Julia's `eigen` orders general complex spectra by (real, imag)
    (LinearAlgebra.eigsortby, stdlib - not in the reference repo)."""
    p = np.lexsort((w.imag, w.real))
    return w[p], v[:, p]


def lu_factorizer(C):
    """Default `factorizer=lu` (src/feast.jl:5): dense -> zgetrf, sparse -> a
    sparse LU (UMFPACK upstream, SuperLU here)."""
    if _is_sparse(C):
        return spla.splu(sp.csc_matrix(C, dtype=complex))
    return sla.lu_factor(np.asarray(C, dtype=complex), check_finite=False)


def left_divide(F, R):
    """Default `left_divider=ldiv!` (src/feast.jl:5): Y = F \\ R."""
    if isinstance(F, tuple):
        return sla.lu_solve(F, R, check_finite=False)
    return F.solve(np.ascontiguousarray(R))


def linsolve(C, R, factorizer=lu_factorizer, left_divider=left_divide):
    """src/utils.jl:175-179."""
    return left_divider(factorizer(C), R)


def normalize_cols(X):
    """src/utils.jl:144-149."""
    X /= np.linalg.norm(X, axis=0)
    return X


def update_R_linear(X, Lam, A, B=None):
    """src/utils.jl:111-116: normalise each column, then R_j = (A - lam_j B) x_j."""
    X /= np.linalg.norm(X, axis=0)
    AX = A @ X
    BX = X if B is None else B @ X
    return np.asarray(AX - BX * Lam[None, :])


def residuals_linear(R):
    """src/utils.jl:166-171: absolute 2-norms (the `A` argument is unused upstream)."""
    return np.linalg.norm(R, axis=0)


def update_R_nep(X, Lam, T):
    """src/utils.jl:104-109: R_j = T(lam_j) x_j after normalising x_j."""
    X /= np.linalg.norm(X, axis=0)
    R = np.empty_like(X)
    for j in range(X.shape[1]):
        R[:, j] = T(Lam[j]) @ X[:, j]
    return R


def _fro(M):
    return spla.norm(M) if _is_sparse(M) else np.linalg.norm(M)


def residuals_nep(R, Lam, T):
    """src/utils.jl:151-157: ||R_j||_2 / ||T(lam_j)||_F."""
    return np.array([np.linalg.norm(R[:, j]) / _fro(T(Lam[j])) for j in range(R.shape[1])])


def beyn_svd_step(Q0, Q1):
    """src/utils.jl:69-77: thin SVD of Q0, A = U' Q1 V S^-1, eig(A), X = U vecs."""
    U, S, Vh = sla.svd(Q0, full_matrices=False, check_finite=False)
    Am = (U.conj().T @ Q1) @ Vh.conj().T
    Am = Am * (1.0 / S)[None, :]
    w, v = sla.eig(Am, check_finite=False)
    w, v = _julia_eig_sort(w, v)
    return w, U @ v


class Timers(dict):
    def add(self, key, t0):
        self[key] = self.get(key, 0.0) + (time.perf_counter() - t0)


# --------------------------------------------------------------------------
# linear drivers  (src/feast.jl)
# --------------------------------------------------------------------------

def _feast_core(X, A, B, contour, iter, eps, store, factorizer, left_divider,
                timers=None, history=None, node_subset=None, reduce_fn=None, col_slice=None):
    """Shared skeleton of feast! (src/feast.jl:10-80) and gen_feast!
    (src/feast.jl:89-156).  B=None is the standard problem (B = I).

    `node_subset` / `reduce_fn` exist so tests can emulate node sharding over
    ranks (each rank accumulates its nodes, then reduce_fn sums Q across ranks);
    `col_slice` = (j0, j1) emulates the column sharding of the device Krylov path
    (a rank runs ALL nodes on the right-hand-side columns j0 .. j1-1 only).
    """
    N, m0 = X.shape
    if A.shape[0] != A.shape[1]:
        raise ValueError("Incorrect dimensions of A, must be square")  # feast.jl:13
    if A.shape[0] != N:
        raise ValueError("Incorrect dimensions of X, must match A")  # feast.jl:15
    tm = timers if timers is not None else Timers()
    nodes = len(contour.nodes)
    my_nodes = range(nodes) if node_subset is None else node_subset
    sparse = _is_sparse(A)
    Ac = A.astype(complex) if sparse else np.asarray(A, dtype=complex)
    if B is None:
        Bop = sp.identity(N, dtype=complex, format="csc") if sparse else np.eye(N, dtype=complex)
    else:
        Bop = B.astype(complex) if _is_sparse(B) else np.asarray(B, dtype=complex)
        if sparse and not _is_sparse(Bop):
            Bop = sp.csc_matrix(Bop)

    def shifted(i):  # A - z_i B   (feast.jl:64,141)
        return Ac - Bop * contour.nodes[i]

    facts = None
    if store:  # feast.jl:28-38 / 107-115
        t0 = time.perf_counter()
        facts = {i: factorizer(shifted(i)) for i in my_nodes}
        tm.add("factor", t0)

    Q = np.array(X, dtype=complex)
    Lam = np.zeros(m0, complex)
    res = np.zeros(m0)
    nit_done = 0
    for nit in range(iter + 1):  # for nit=0:iter
        nit_done = nit
        t0 = time.perf_counter()
        Q, _ = np.linalg.qr(Q)  # feast.jl:41
        tm.add("qr", t0)
        t0 = time.perf_counter()
        R = A @ Q
        Aq = Q.conj().T @ R  # feast.jl:42-43
        if B is not None:
            R = B @ Q
            Bq = Q.conj().T @ R  # feast.jl:120-121
        tm.add("project", t0)
        t0 = time.perf_counter()
        if B is None:
            w, v = sla.eig(Aq, check_finite=False)  # feast.jl:45
        else:
            w, v = sla.eig(Aq, Bq, check_finite=False)  # feast.jl:122
        Lam, Xq = _julia_eig_sort(w, v)
        tm.add("reduced_eig", t0)
        t0 = time.perf_counter()
        X[:, :] = Q @ Xq  # feast.jl:48
        R = update_R_linear(X, Lam, A, B)  # feast.jl:49 / 126
        res = residuals_linear(R)  # feast.jl:50
        tm.add("recover_residual", t0)
        inside = in_contour(Lam, contour)
        if history is not None:
            history.append((nit, int(inside.sum()), float(res[inside].max()) if inside.any() else np.nan))
        if inside.any() and res[inside].max() < eps:  # feast.jl:53
            break
        if nit < iter:  # feast.jl:57
            Q = np.zeros((N, m0), complex)
            for i in my_nodes:
                resolvent = 1.0 / (contour.nodes[i] - Lam)  # feast.jl:60
                if store:
                    t0 = time.perf_counter()
                    temp = left_divider(facts[i], R)
                    tm.add("solve", t0)
                else:
                    t0 = time.perf_counter()
                    F = factorizer(shifted(i))
                    tm.add("factor", t0)
                    t0 = time.perf_counter()
                    temp = left_divider(F, R)
                    tm.add("solve", t0)
                t0 = time.perf_counter()
                temp = X - temp  # feast.jl:68
                temp *= (resolvent * contour.weights[i])[None, :]  # feast.jl:69
                if col_slice is not None:   # the other columns belong to other ranks
                    temp[:, :col_slice[0]] = 0.0
                    temp[:, col_slice[1]:] = 0.0
                Q += temp  # feast.jl:70
                tm.add("accumulate", t0)
            if reduce_fn is not None:
                Q = reduce_fn(Q)
    inside = in_contour(Lam, contour)
    info = {"iterations": nit_done, "Lam_all": Lam, "res_all": res, "inside": inside}
    return Lam[inside], X[:, inside], res[inside], info


def feast(X, A, contour=None, *, nodes=8, iter=10, c=0.0 + 0.0j, r=1.0, eps=1e-12,
          store=False, factorizer=lu_factorizer, left_divider=left_divide,
          timers=None, history=None, full=False, **kw):
    """feast!(X, A; ...) / feast!(X, A, contour; ...)  -- src/feast.jl:3-80.
    X is mutated in place (all m0 normalised Ritz vectors)."""
    if contour is None:
        contour = circular_contour_trapezoidal(c, r, nodes)  # feast.jl:6
    out = _feast_core(X, A, None, contour, iter, eps, store, factorizer, left_divider,
                      timers, history, **kw)
    return out if full else out[:3]


def gen_feast(X, A, B, contour=None, *, nodes=8, iter=10, c=0.0 + 0.0j, r=1.0, eps=1e-12,
              store=False, factorizer=lu_factorizer, left_divider=left_divide,
              timers=None, history=None, full=False, **kw):
    """gen_feast!(X, A, B[, contour]; ...) -- src/feast.jl:82-156."""
    if contour is None:
        contour = circular_contour_trapezoidal(c, r, nodes)  # feast.jl:85
        store = False  # the convenience wrapper drops `store` (feast.jl:86)
    out = _feast_core(X, A, B, contour, iter, eps, store, factorizer, left_divider,
                      timers, history, **kw)
    return out if full else out[:3]


def dual_gen_feast(Xr, Xl, A, B, contour=None, *, nodes=8, iter=10, c=0.0 + 0.0j, r=1.0,
                   eps=1e-12, store=False, full=False):
    """dual_gen_feast! -- src/feast.jl:158-257 (two-sided, non-Hermitian pencils).

    Two upstream defects are restated with their INTENDED semantics (the only
    upstream call with an assertion is commented out, test/runtests.jl:24-26, so
    this routine is PARITY UNPINNED):
      * `Diagonal(1.0/S.S)` (feast.jl:200-201) is Number/Vector in Julia, not an
        elementwise inverse.  The comment at feast.jl:205 ("Bq = Q' * Q = I") shows
        the intent: bi-orthonormalise so that Ql' B Qr = I, i.e. scale BOTH bases
        by the elementwise Sigma^-1/2 (Sigma^-1 on both sides would give Sigma^-1
        and overflows once the surplus directions have been filtered away).  The
        reduced pencil (Aq, Bq) is formed explicitly, so any nonsingular scaling
        gives the same Ritz values.
      * `update_R!(Xl, Rl, L, A', B')` (feast.jl:214) uses L where a left
        eigenpair satisfies A'y = conj(l) B'y; the conj resolvent at feast.jl:238
        is only consistent with R_l = (A' - conj(l) B') y, which is what is used.
    """
    if contour is None:
        contour = circular_contour_trapezoidal(c, r, nodes)
    N, m0 = Xl.shape
    if A.shape[0] != A.shape[1]:
        raise ValueError("Incorrect dimensions of A, must be square")
    if A.shape[0] != N:
        raise ValueError("Incorrect dimensions of X, must match A")
    sparse = _is_sparse(A)
    Ac = A.astype(complex) if sparse else np.asarray(A, dtype=complex)
    if B is None:
        Bop = sp.identity(N, dtype=complex, format="csc") if sparse else np.eye(N, dtype=complex)
    else:
        Bop = B.astype(complex) if _is_sparse(B) else np.asarray(B, dtype=complex)
    AH, BH = Ac.conj().T, Bop.conj().T
    nn = len(contour.nodes)
    Qr, Ql = np.array(Xr, dtype=complex), np.array(Xl, dtype=complex)
    rf = lf = None
    if store:  # feast.jl:181-196
        rf = [lu_factorizer(Ac - Bop * z) for z in contour.nodes]
        lf = [lu_factorizer((Ac - Bop * z).conj().T) for z in contour.nodes]
    Lam = np.zeros(m0, complex)
    resr = np.zeros(m0)
    nit_done = 0
    for nit in range(iter + 1):
        nit_done = nit
        U, S, Vh = sla.svd(Ql.conj().T @ (Bop @ Qr))  # feast.jl:199
        sc = 1.0 / np.sqrt(np.maximum(S, S[0] * 1e-280))  # Sigma^-1/2 on both sides: Ql' B Qr = I (feast.jl:205 comment)
        Qr = Qr @ Vh.conj().T * sc[None, :]
        Ql = Ql @ U * sc[None, :]
        Aq = Ql.conj().T @ (Ac @ Qr)
        Bq = Ql.conj().T @ (Bop @ Qr)
        w, v = sla.eig(Aq, Bq)  # feast.jl:206
        Lam, Xq = _julia_eig_sort(w, v)
        Xr[:, :] = Qr @ Xq
        # Left vectors: eigen(Aq', Bq') has eigenvalues conj(Lam); Julia sorts them by
        # (re, im) of conj(Lam) which permutes conjugate pairs relative to Lam
        # (feast.jl:210-212).  Pair them with Lam explicitly (intended semantics).
        wl, vl = sla.eig(Aq.conj().T, Bq.conj().T)
        order = [int(np.argmin(np.abs(wl - np.conj(l)))) for l in Lam]
        Xl[:, :] = Ql @ vl[:, order]
        Rr = update_R_linear(Xr, Lam, Ac, Bop)  # feast.jl:213
        Rl = update_R_linear(Xl, np.conj(Lam), AH, BH)  # feast.jl:214 (see docstring)
        resr = residuals_linear(Rr)
        inside = in_contour(Lam, contour)
        if inside.any() and resr[inside].max() < eps:
            break
        if nit < iter:
            Qr = np.zeros((N, m0), complex)
            Ql = np.zeros((N, m0), complex)
            for i in range(nn):
                z, wgt = contour.nodes[i], contour.weights[i]
                temp = left_divide(rf[i], Rr) if store else linsolve(Ac - Bop * z, Rr)
                Qr += (Xr - temp) * ((1.0 / (z - Lam)) * wgt)[None, :]
                temp = left_divide(lf[i], Rl) if store else linsolve((Ac - Bop * z).conj().T, Rl)
                Ql += (Xl - temp) * ((1.0 / (np.conj(z) - np.conj(Lam))) * np.conj(wgt))[None, :]
    inside = in_contour(Lam, contour)
    info = {"iterations": nit_done, "Lam_all": Lam, "res_all": resr, "inside": inside}
    out = (Lam[inside], Xr[:, inside], Xl[:, inside], resr[inside])
    return out + (info,) if full else out


# --------------------------------------------------------------------------
# nonlinear driver  (src/nlfeast.jl:2-84)
# --------------------------------------------------------------------------

def nlfeast(T, X, nodes, iter, *, c=0.0 + 0.0j, r=1.0, eps=10e-12, store=True,
            spurious=1e-5, factorizer=lu_factorizer, left_divider=left_divide,
            timers=None, history=None, node_subset=None, reduce_fn=None):
    """nlfeast!(T, X, nodes, iter; ...) -- returns (Lam, X, res), all m0 entries,
    NOT filtered by the contour (nlfeast.jl:83).  X mutated in place."""
    N, m0 = X.shape
    tm = timers if timers is not None else Timers()
    theta = np.linspace(np.pi / nodes, 2 * np.pi - np.pi / nodes, nodes)  # nlfeast.jl:8
    zs = r * np.exp(1j * theta) + c
    ws = r * np.exp(1j * theta) / nodes
    my_nodes = range(nodes) if node_subset is None else node_subset
    Xq, _ = np.linalg.qr(X)  # nlfeast.jl:12-13
    X[:, :] = Xq
    facts = None
    if store:  # nlfeast.jl:17-28
        t0 = time.perf_counter()
        facts = {i: factorizer(T(zs[i])) for i in my_nodes}
        tm.add("factor", t0)
    Lam = np.zeros(m0, complex)
    res = np.zeros(m0)
    R = None
    for nit in range(iter + 1):
        Q0 = np.zeros((N, m0), complex)
        Q1 = np.zeros((N, m0), complex)
        for i in my_nodes:  # nlfeast.jl:36-61
            rhs = X if nit == 0 else R
            t0 = time.perf_counter()
            if store:
                Y = left_divider(facts[i], rhs)
            else:
                Y = left_divider(factorizer(T(zs[i])), rhs)
            tm.add("solve", t0)
            t0 = time.perf_counter()
            if nit == 0:
                Tinv = Y * ws[i]  # nlfeast.jl:39-45
            else:
                Tinv = (X - Y) * ((1.0 / (zs[i] - Lam)) * ws[i])[None, :]  # :46-55
            Q0 += Tinv
            Q1 += Tinv * zs[i]
            tm.add("accumulate", t0)
        if reduce_fn is not None:
            Q0, Q1 = reduce_fn(Q0), reduce_fn(Q1)
        t0 = time.perf_counter()
        Lam, Xn = beyn_svd_step(Q0, Q1)  # nlfeast.jl:64
        X[:, :] = Xn
        tm.add("beyn", t0)
        t0 = time.perf_counter()
        R = update_R_nep(X, Lam, T)  # nlfeast.jl:66
        res = residuals_nep(R, Lam, T)  # nlfeast.jl:67
        tm.add("recover_residual", t0)
        inside = in_contour(Lam, c, r)
        res_inside = res[inside]
        if history is not None:
            history.append((nit, int(inside.sum()), float(res_inside.max()) if inside.any() else np.nan))
        if res_inside.size > 0 and res_inside.max() < eps:  # nlfeast.jl:73
            break
        good = res_inside[res_inside < spurious]
        if nit > 1 and good.size > 0 and good.max() < eps:  # nlfeast.jl:76
            break
    normalize_cols(X)  # nlfeast.jl:82
    return Lam, X, res


# --------------------------------------------------------------------------
# one-shot contour solvers and higher moments  (src/beyn.jl, src/nlfeast.jl:173-318)
# --------------------------------------------------------------------------

def beyn(T, A, X, nodes, *, c=0.0 + 0.0j, r=1.0):
    """beyn(T, A, X, nodes; c, r)  (src/beyn.jl:2-34): Beyn's integral method, one contour pass, two moments.
    A only supplies the dimensions (:5-9).  Weights exp(i theta)/nodes (no factor r, :19-20: it cancels in
    U' Q1 V S^-1).  Returns (L[p], X[:, p], res[p]) sorted by the ABSOLUTE residual ||T(l) x|| (:29-33)."""
    N, m0 = X.shape
    if A.shape[0] != A.shape[1]:
        raise ValueError("Incorrect dimensions of A, must be square")
    if A.shape[0] != N:
        raise ValueError("Incorrect dimensions of X0, must match A")
    theta = np.linspace(np.pi / nodes, 2 * np.pi - np.pi / nodes, nodes)
    Q0 = np.zeros((N, m0), complex)
    Q1 = np.zeros((N, m0), complex)
    for th in theta:
        z = r * np.exp(1j * th) + c
        temp = left_divide(lu_factorizer(T(z)), X)
        Q0 += temp * (np.exp(1j * th) / nodes)
        Q1 += temp * (z * np.exp(1j * th) / nodes)
    Lam, Xn = beyn_svd_step(Q0, Q1)
    res = np.array([np.linalg.norm(T(Lam[i]) @ Xn[:, i]) for i in range(m0)])
    p = np.argsort(res, kind="stable")
    return Lam[p], Xn[:, p], res[p]


def contour_moments(T, X, nodes, nmom, c, r, rhs=None, Lam=None):
    """S_p = sum_k z_k^p (T(z_k)^-1 X) w_k, w_k = r exp(i theta_k)/nodes  (src/beyn.jl:50-56, src/nlfeast.jl:195-213);
    with rhs/Lam given: the residual-inverse-iteration form (X - T(z_k)^-1 R) diag(w_k/(z_k - lam))  (nlfeast.jl:255-275)."""
    N, m0 = X.shape
    theta = np.linspace(np.pi / nodes, 2 * np.pi - np.pi / nodes, nodes)
    S = [np.zeros((N, m0), complex) for _ in range(nmom)]
    for th in theta:
        z = r * np.exp(1j * th) + c
        w = r * np.exp(1j * th) / nodes
        F = lu_factorizer(T(z))
        if rhs is None:
            temp = left_divide(F, X) * w
        else:
            temp = (X - left_divide(F, rhs)) * ((1.0 / (z - Lam)) * w)[None, :]
        for p in range(nmom):
            S[p] += temp * z ** p
    return S


def block_SS(T, X, nodes=16, moments=2, *, c=0.0 + 0.0j, r=1.0, Y=None, rng=None):
    """block_SS!(T, X, nodes, moments; c, r)  (src/beyn.jl:36-94): block Sakurai-Sugiura with a probe block Y
    (upstream: rand(ComplexF64, N, m0), unseeded -- pass Y for reproducibility).  Returns (L, X, res) with
    n <= moments*m0 entries (numerical rank of the Hankel matrix at 1e-13, :78), res relative (:92)."""
    N, m0 = X.shape
    K = moments * m0
    Xo, _ = np.linalg.qr(X)
    if Y is None:
        rng = np.random.default_rng(0) if rng is None else rng
        Y = rng.random((N, m0)) + 1j * rng.random((N, m0))
    l = Y.shape[1]
    S = contour_moments(T, Xo, nodes, 2 * moments + 1, c, r)
    Q0 = np.zeros((l * moments, K), complex)
    Q1 = np.zeros((l * moments, K), complex)
    for i in range(1, moments + 1):
        for j in range(1, moments + 1):
            Q0[(i - 1) * l:i * l, (j - 1) * m0:j * m0] = Y.conj().T @ S[i + j - 1]     # :66
            Q1[(i - 1) * l:i * l, (j - 1) * m0:j * m0] = Y.conj().T @ S[i + j]         # :67
    U, sv, Vh = sla.svd(Q0, full_matrices=False, check_finite=False)
    n = min(int(np.count_nonzero(sv / sv[0] > 1e-13)), K)
    V = Vh.conj().T
    H1 = U[:, :n].conj().T @ Q1 @ V[:, :n]
    H0 = U[:, :n].conj().T @ Q0 @ V[:, :n]
    Lam, Xq = sla.eig(H1, H0, check_finite=False)
    Xn = np.hstack(S[:moments]) @ V[:, :n] @ Xq                                        # :87
    Xn /= np.linalg.norm(Xn, axis=0)
    res = np.array([np.linalg.norm(T(Lam[i]) @ Xn[:, i]) / _fro(T(Lam[i])) for i in range(n)])
    return Lam, Xn, res


def _moments_reduce(S, moments, N, m0):
    """Block-Hankel Q0, Q1 (moments*N x moments*m0) from S_0 .. S_{2 moments - 1}, tall SVD, A = U' Q1 V S^-1, eig,
    Y = U[1:N, :] vecs  (src/nlfeast.jl:216-231)."""
    K = moments * m0
    Q0 = np.zeros((moments * N, K), complex)
    Q1 = np.zeros((moments * N, K), complex)
    for i in range(1, moments + 1):
        for j in range(1, moments + 1):
            Q0[(i - 1) * N:i * N, (j - 1) * m0:j * m0] = S[i + j - 2]
            Q1[(i - 1) * N:i * N, (j - 1) * m0:j * m0] = S[i + j - 1]
    U, sv, Vh = sla.svd(Q0, full_matrices=False, check_finite=False)
    Am = (U.conj().T @ Q1) @ Vh.conj().T * (1.0 / sv)[None, :]
    w, v = sla.eig(Am, check_finite=False)
    w, v = _julia_eig_sort(w, v)
    return w, U[:N, :] @ v


def _update_R_moments(Y, Lam, T):
    """src/utils.jl:118-134: normalise, R_i = T(l_i) y_i, relative residuals, everything sorted by residual."""
    R = update_R_nep(Y, Lam, T)
    res = residuals_nep(R, Lam, T)
    p = np.argsort(res, kind="stable")
    return Y[:, p], R[:, p], Lam[p], res[p]


def nlfeast_moments(T, X, nodes, iter, *, c=0.0 + 0.0j, r=1.0, eps=10e-12, moments=2, store=True, spurious=1e-5,
                    history=None):
    """nlfeast_moments!(T, X, nodes, iter; ...)  (src/nlfeast.jl:173-318): nlfeast with 2*moments moment accumulators and
    a moments*m0-dimensional Beyn reduction per pass.  Returns (L, Y, res) with moments*m0 entries sorted by residual;
    X is overwritten by the m0 best (unit-norm) vectors."""
    N, m0 = X.shape
    S = contour_moments(T, X, nodes, 2 * moments, c, r)                    # :195-213 (X is NOT orthonormalised here)
    Lam, Y = _moments_reduce(S, moments, N, m0)
    Y, R, Lam, res = _update_R_moments(Y, Lam, T)                          # :236
    X[:, :] = Y[:, :m0]
    for nit in range(1, iter + 1):
        S = contour_moments(T, X, nodes, 2 * moments, c, r, rhs=R[:, :m0], Lam=Lam[:m0])   # :255-275
        Lam, Y = _moments_reduce(S, moments, N, m0)
        Y, R, Lam, res = _update_R_moments(Y, Lam, T)
        X[:, :] = Y[:, :m0]
        inside = in_contour(Lam[:m0], c, r)
        res_inside = res[:m0][inside]
        if history is not None:
            history.append((nit, int(inside.sum()), float(res_inside.max()) if inside.any() else np.nan))
        if res_inside.size > 0 and res_inside.max() < eps:                 # :297
            break
        good = res_inside[res_inside < spurious]
        if nit > 1 and good.size > 0 and good.max() < eps:                 # :300
            break
    normalize_cols(X)
    return Lam, Y, res


# --------------------------------------------------------------------------
# inexact-inner-solve precedents  (src/feast_experimental.jl:1-60, src/nlfeast.jl:87-171)
# --------------------------------------------------------------------------

def _bicgstab_cols(Z, Bm, tol, X0=None, maxiter=None):
    """Column-by-column Krylov solves as in `for m=1:m0 temp[:, m] .= bicgstabl(ZmA, X[:, m])`
    (feast_experimental.jl:27-29).  IterativeSolvers.bicgstabl is restated with scipy's BiCGStab
    (same Krylov family; only the achieved tolerance matters to the outer iteration)."""
    import scipy.sparse.linalg as spla
    Bm = np.asarray(Bm, dtype=complex)
    out = np.zeros_like(Bm)
    for j in range(Bm.shape[1]):
        x0 = None if X0 is None else np.ascontiguousarray(X0[:, j])
        x, info = spla.bicgstab(Z, Bm[:, j], x0=x0, rtol=tol, atol=0.0, maxiter=maxiter)
        out[:, j] = x
    return out


def ifeast(A, X0, nodes, iter, *, c=0.0 + 0.0j, r=1.0, eps=0.05, tol=None):
    """ifeast!(A, X0, nodes, iter; c, r, debug, eps)  (src/feast_experimental.jl:1-60): plain (non-RII) FEAST with
    inexact per-column Krylov solves of (zI - A) Y = X, `iter` passes, NO orthonormalisation (the reduced problem is
    generalized: Aq = Q'AQ, Bq = Q'Q), returns ALL m0 Ritz pairs with absolute residuals ||A x - x l||."""
    N, m0 = X0.shape
    if A.shape[0] != A.shape[1]:
        raise ValueError("Incorrect dimensions of A, must be square")
    if A.shape[0] != N:
        raise ValueError("Incorrect dimensions of X0, must match A")
    tol = np.sqrt(np.finfo(float).eps) if tol is None else tol   # IterativeSolvers' default reltol
    X = np.array(X0, dtype=complex)
    theta = np.linspace(np.pi / nodes, 2 * np.pi - np.pi / nodes, nodes)
    Lam = np.zeros(m0, complex)
    res = np.zeros(m0)
    eye = sp.identity(N, format="csr", dtype=complex) if _is_sparse(A) else np.eye(N, dtype=complex)
    for _ in range(iter):
        Q = np.zeros((N, m0), complex)
        for i in range(nodes):
            z = r * np.exp(1j * theta[i]) + c
            ZmA = eye * z - A
            Q += _bicgstab_cols(ZmA, X, tol) * (np.exp(1j * theta[i]) / nodes)     # :31 (no factor r upstream)
        Aq = Q.conj().T @ (A @ Q)
        Bq = Q.conj().T @ Q
        w, v = sla.eig(Aq, Bq, check_finite=False)
        Lam, v = _julia_eig_sort(w, v)
        X = Q @ v
        X = X / np.linalg.norm(X, axis=0)[None, :]
        res = np.linalg.norm(A @ X - X * Lam[None, :], axis=0)
    return Lam, X, res


def nlfeast_it(T, X, nodes, iter, *, c=0.0 + 0.0j, r=1.0, eps=0.05):
    """nlfeast_it!(T, X, nodes, iter; c, r, debug, eps)  (src/nlfeast.jl:87-171): nlfeast with per-column bicgstabl
    solves, tolerance 1e-3 in the first pass (:106) and 1e-8 warm-started from the previous solution afterwards
    (:139); stops when max(res[inside]) < eps (:164); returns all m0 pairs, relative residuals."""
    N, m0 = X.shape
    theta = np.linspace(np.pi / nodes, 2 * np.pi - np.pi / nodes, nodes)
    zs = r * np.exp(1j * theta) + c
    ws = r * np.exp(1j * theta) / nodes
    Tz = [T(z) for z in zs]
    Tinv = [None] * nodes
    Q0 = np.zeros((N, m0), complex)
    Q1 = np.zeros((N, m0), complex)
    for i in range(nodes):
        Tinv[i] = _bicgstab_cols(Tz[i], X, 1e-3)
        Q0 += Tinv[i] * ws[i]
        Q1 += Tinv[i] * ws[i] * zs[i]
    Lam, Xn = beyn_svd_step(Q0, Q1)
    X[:, :] = Xn
    R = update_R_nep(X, Lam, T)
    res = residuals_nep(R, Lam, T)
    for nit in range(1, iter + 1):
        Q0[:] = 0
        Q1[:] = 0
        for i in range(nodes):
            Tinv[i] = _bicgstab_cols(Tz[i], R, 1e-8, X0=Tinv[i])
            Tm = (X - Tinv[i]) * (ws[i] / (zs[i] - Lam))[None, :]
            Q0 += Tm
            Q1 += Tm * zs[i]
        Lam, Xn = beyn_svd_step(Q0, Q1)
        X[:, :] = Xn
        R = update_R_nep(X, Lam, T)
        res = residuals_nep(R, Lam, T)
        inside = np.abs(Lam - c) <= r
        if inside.any() and res[inside].max() < eps:
            break
    X[:, :] = normalize_cols(X)
    return Lam, X, residuals_nep(update_R_nep(X, Lam, T), Lam, T)


def polynomial(coeffs):
    """T(z) = sum_i z^i A_i, the closures of test/butterfly.jl:61, test/polynomial.jl:9-11."""
    def T(z):
        acc = coeffs[-1] * z
        for Ai in coeffs[-2:0:-1]:
            acc = (acc + Ai) * z
        return acc + coeffs[0]
    return T


def companion(coeffs):
    """src/companion.jl:1-28: dense companion linearisation, returns all N*L
    eigenpairs with relative residuals (the reference's exact-answer generator)."""
    A = [np.asarray(a.todense() if _is_sparse(a) else a, dtype=complex) for a in coeffs]
    N = A[0].shape[0]
    L = len(A) - 1
    C1 = np.zeros((N * L, N * L), complex)
    C2 = np.zeros((N * L, N * L), complex)
    C1[:N, :N] = A[0]
    for i in range(N, N * L):
        C1[i, i] = 1
        C2[i, i - N] = 1
    for i in range(L):
        C2[:N, N * i:N * (i + 1)] = -A[i + 1]
    w, V = sla.eig(C1, C2)
    w, V = _julia_eig_sort(w, V)
    X = V[(L - 1) * N:L * N, :].copy()
    res = np.zeros(N * L)
    for i in range(N * L):
        X[:, i] /= np.linalg.norm(X[:, i])
        Rv = sum(A[j] @ X[:, i] * w[i] ** j for j in range(L + 1))
        res[i] = np.linalg.norm(Rv) / np.linalg.norm(sum(A[j] * w[i] ** j for j in range(L + 1)))
    return w, X, res


def contour_estimate_eig(A, contour, B=None, *, samples=None, rng=None, X=None):
    """src/stochastic.jl:2-33: Hutchinson estimate of the eigenvalue count,
    sum_i w_i tr(X' (z_i B - A)^-1 X) / samples (note the z B - A sign, :24)."""
    N = A.shape[0]
    m0 = min(100, N) if samples is None else samples
    if X is None:
        rng = np.random.default_rng(0) if rng is None else rng
        X = (rng.standard_normal((N, m0)) + 1j * rng.standard_normal((N, m0))) / np.sqrt(2)
    m0 = X.shape[1]  # `samples` is the number of probe columns (stochastic.jl:7,27)
    sparse = _is_sparse(A)
    Ac = A.astype(complex) if sparse else np.asarray(A, dtype=complex)
    if B is None:
        Bop = sp.identity(N, dtype=complex, format="csc") if sparse else np.eye(N, dtype=complex)
    else:
        Bop = B.astype(complex) if _is_sparse(B) else np.asarray(B, dtype=complex)
    est = 0.0
    for z, w in zip(contour.nodes, contour.weights):
        temp = linsolve(Bop * z - Ac, X)
        est += np.trace(X.conj().T @ temp) * w / m0
    return float(np.real(est))


# --------------------------------------------------------------------------
# MatrixMarket reader for data/*.mtx  (test/polynomial.jl:5-7 uses MatrixMarket.jl)
# --------------------------------------------------------------------------

def mmread(path):
    with open(path) as f:
        header = f.readline().split()
        field_t, sym = header[3], header[4]
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        nr, nc, nnz = (int(t) for t in line.split())
        rows = np.empty(nnz, np.int64)
        cols = np.empty(nnz, np.int64)
        vals = np.empty(nnz, complex if field_t == "complex" else float)
        for k in range(nnz):
            t = f.readline().split()
            rows[k], cols[k] = int(t[0]) - 1, int(t[1]) - 1
            vals[k] = complex(float(t[2]), float(t[3])) if field_t == "complex" else float(t[2])
    M = sp.coo_matrix((vals, (rows, cols)), shape=(nr, nc)).tocsc()
    if sym == "symmetric":
        M = M + sp.triu(M.T, 1)
    return M
