"""numpy restatement of the multigrid V-cycle and preconditioned COCG of csrc/amg.cu / krylov.cu (test infrastructure):
reads the hierarchy built by the C++ setup through the host-only debug ABI and runs the same cycle in numpy."""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from feastsolver_jl_b200 import _lib


def union_values(slots):
    S = [sp.csr_matrix(s) for s in slots]
    U = sum(abs(s) for s in S).tocsr()
    U.sort_indices()
    n = U.shape[0]
    rows = np.repeat(np.arange(n), np.diff(U.indptr)).astype(np.int64)
    lin = rows * n + U.indices
    vals = np.zeros((len(S), U.nnz))
    for k, s in enumerate(S):
        co = s.tocoo()
        vals[k, np.searchsorted(lin, co.row.astype(np.int64) * n + co.col)] = co.data
    return U.indptr.astype(np.int64), U.indices.astype(np.int32), vals


def cpp_hierarchy(slots, max_coarse=4096):
    lib = _lib.load()
    rp, ci, vals = union_values(slots)
    n, ns = rp.size - 1, vals.shape[0]
    nl, secs = C.c_int(0), C.c_double(0)
    h = lib.feast_debug_amg_build(n, _lib.ptr(rp), _lib.ptr(ci), ns, _lib.ptr(vals), max_coarse, C.byref(nl), C.byref(secs))
    levels = []
    for l in range(nl.value):
        nn, nnz, nc, pnnz, rho = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_double()
        lib.feast_debug_amg_level_info(h, l, C.byref(nn), C.byref(nnz), C.byref(nc), C.byref(pnnz), C.byref(rho))
        rp_, ci_, v_ = np.zeros(nn.value + 1, np.int32), np.zeros(nnz.value, np.int32), np.zeros((ns, nnz.value))
        prp, pci, pv = np.zeros(nn.value + 1, np.int32), np.zeros(max(pnnz.value, 1), np.int32), np.zeros(max(pnnz.value, 1))
        lib.feast_debug_amg_level_get(h, l, _lib.ptr(rp_), _lib.ptr(ci_), _lib.ptr(v_), _lib.ptr(prp), _lib.ptr(pci), _lib.ptr(pv))
        L = {"n": nn.value, "rho": rho.value,
             "slots": [sp.csr_matrix((v_[k], ci_, rp_), shape=(nn.value, nn.value)) for k in range(ns)]}
        if nc.value:
            L["P"] = sp.csr_matrix((pv[:pnnz.value], pci[:pnnz.value], prp), shape=(nn.value, nc.value))
        levels.append(L)
    lib.feast_debug_amg_free(h)
    return levels, secs.value


class VCycle:
    """y = M^-1 r: V(1,1) with damped Jacobi, w = 2 / (1.1 rho + rho / 30), exact coarsest solve."""

    def __init__(self, levels, coefs):
        self.L = levels
        self.Z = [sum(c * s for c, s in zip(coefs, l["slots"])).tocsr() for l in levels]
        self.Zinv = np.linalg.inv(self.Z[-1].toarray())

    def cycle(self, l, r):
        if l == len(self.L) - 1:
            return self.Zinv @ r
        Z, P = self.Z[l], self.L[l]["P"]
        w = 2.0 / (1.1 * self.L[l]["rho"] + self.L[l]["rho"] / 30.0)
        dinv = (w / Z.diagonal())[:, None]
        y = dinv * r
        y = y + P @ self.cycle(l + 1, P.T @ (r - Z @ y))
        return y + dinv * (r - Z @ y)

    def __call__(self, r):
        return self.cycle(0, r)


def pcocg(Z, b, tol, maxit, prec=None):
    x = np.zeros_like(b)
    r = b.copy()
    z = prec(r) if prec else r
    p = z.copy()
    rho = np.sum(r * z, axis=0)
    bn = np.linalg.norm(b, axis=0)
    for it in range(1, maxit + 1):
        q = Z @ p
        al = rho / np.sum(p * q, axis=0)
        x += p * al
        r -= q * al
        rel = np.linalg.norm(r, axis=0) / bn
        if rel.max() < tol:
            break
        z = prec(r) if prec else r
        rho2 = np.sum(r * z, axis=0)
        p = z + p * (rho2 / rho)
        rho = rho2
    return x, it, rel.max()
